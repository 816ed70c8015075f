// channelize_iq — command-line companion of the reference's record tools: takes a recording written
// by blade_record_iq_* / usrp_record_iq_* (cpp/IqPacket.h header + interleaved I/Q), channelizes it on
// the GPU through libchannelizer and writes the channel matrix and the PDW table.  Given a DIRECTORY it
// watches it and processes every dwell file the recorders drop there (one file per dwell,
// cpp/blade_record_iq_12bit.cpp:316-324), in name order -- the recorder's file names are UTC time stamps
// (cpp/Helper.cpp:6-23), so name order is time order.
// Style follows the reference tools: positional arguments, a usage text, `return __LINE__` on failure
// (cpp/blade_record_iq_12bit.cpp:31-37,54-59).
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "channelizer.h"

#define CHECK(call)                                                                              \
  do {                                                                                           \
    const int status_ = (call);                                                                  \
    if (status_ != 0) {                                                                          \
      std::cerr << #call << " failed: " << chz_strerror(status_) << " " << chz_last_cuda_error() \
                << std::endl;                                                                    \
      return __LINE__;                                                                           \
    }                                                                                            \
  } while (0)

namespace {

struct Options {
  uint32_t channels = 0, tapsPerBand = 12, oversample = 1;
  double snrThresholdDb = 15.0;
  bool bugCompat = false;   // reproduce create_pdws_channelized.m:114 (phase of column 1 for every bin)
};

// Closes the recording and frees the pinned output buffer on every exit path of processOne.
struct FileGuard {
  chz_iq_t* file = nullptr;
  chz_cf32* out = nullptr;
  ~FileGuard() {
    if (out) chz_free_host(out);
    if (file) chz_close_iq(file);
  }
};

// One channelizer handle is kept across files while the geometry stays the same (chz_reset = a fresh
// dsp.Channelizer per file, matlab/create_pdws_channelized.m:33).
struct Engine {
  chz_t* chan = nullptr;
  uint32_t channels = 0;
  ~Engine() { if (chan) chz_destroy(chan); }
};

// 0 = done, > 0 = failure line, -1 = the file is not complete yet (watch mode retries it; `why` says what is wrong)
int processOne(Engine& eng, const Options& opt, const std::string& path, const std::string& prefix, int* why = nullptr) {
  FileGuard guard;
  chz_iq_t*& file = guard.file;
  chz_iq_info_t info;
  const int rc = chz_open_iq(path.c_str(), &file, &info);
  if (rc == CHZ_ESIZE || rc == CHZ_EIO) {              // header or payload still being written
    if (why) *why = rc;
    return -1;
  }
  CHECK(rc);
  std::cout << path << ": file format " << info.format << ", " << info.num_samples << " samples, " << info.bit_width
            << " bits, fs " << info.fs_sps << " sps, fc " << info.fc_hz << " Hz, board '" << info.board_name << "'"
            << std::endl;
  uint32_t channels = opt.channels;
  if (channels == 0) channels = (uint32_t)(info.fs_sps * 1e-6 + 0.5);   // create_pdws_channelized.m:31
  if (!eng.chan || eng.channels != channels) {
    if (eng.chan) chz_destroy(eng.chan);
    eng.chan = nullptr;
    std::vector<float> taps((size_t)channels * opt.tapsPerBand);
    CHECK(chz_design_prototype(channels, opt.tapsPerBand, 80.0, taps.data()));
    CHECK(chz_create(channels, taps.data(), (uint32_t)taps.size(), opt.oversample, &eng.chan));
    CHECK(chz_set_option(eng.chan, CHZ_OPT_RETAIN, 1));   // chz_pdws below runs over the rows kept on the GPU
    eng.channels = channels;
  } else {
    CHECK(chz_reset(eng.chan));
  }
  chz_t* chan = eng.chan;

  const uint64_t rows = chz_rows_for(chan, info.num_samples);
  chz_cf32*& out = guard.out;
  if (!prefix.empty()) {
    out = (chz_cf32*)chz_alloc_host(rows * channels * sizeof(chz_cf32));
    if (!out) { std::cerr << "pinned allocation failed" << std::endl; return __LINE__; }
  }
  uint64_t got = 0;
  const auto t0 = std::chrono::steady_clock::now();
  CHECK(chz_process(chan, chz_iq_payload(file), info.num_samples, info.bit_width, out, out ? rows : 0, &got));
  const auto t1 = std::chrono::steady_clock::now();

  chz_pdw_params_t prm;
  std::memset(&prm, 0, sizeof prm);
  prm.snr_threshold_db = opt.snrThresholdDb;
  prm.sat_level = 0.9999;
  prm.fc_hz = (double)info.fc_hz;
  prm.fs_sps = (double)info.fs_sps;
  prm.t0 = info.sample_start_time;
  prm.use_trailing_threshold = 0;
  prm.reproduce_phase_bug = opt.bugCompat ? 1 : 0;
  uint64_t npdw = 0;
  const int status = chz_pdws(chan, &prm, nullptr, 0, &npdw);
  if (status != 0 && status != CHZ_ECAPACITY) CHECK(status);
  std::vector<chz_pdw_t> pdws(npdw);
  if (npdw) CHECK(chz_pdws_fetch(chan, pdws.data(), npdw, &npdw));
  const auto t2 = std::chrono::steady_clock::now();

  const double sChan = std::chrono::duration<double>(t1 - t0).count();
  const double sPdw = std::chrono::duration<double>(t2 - t1).count();
  std::cout << "Channelized " << got << " rows x " << channels << " channels in " << sChan << " s ("
            << info.num_samples / sChan * 1e-6 << " MS/s incl. host<->device copies); " << npdw << " PDWs in "
            << sPdw << " s" << std::endl;

  if (!prefix.empty()) {
    const std::string chanName = prefix + ".cf32";
    FILE* fc = std::fopen(chanName.c_str(), "wb");
    if (!fc) return __LINE__;
    std::fwrite(out, sizeof(chz_cf32), got * channels, fc);
    std::fclose(fc);
    // the table is written under a temporary name and renamed, so a consumer never sees half of it
    const std::string pdwName = prefix + ".pdw.csv", tmpName = pdwName + ".part";
    FILE* fp = std::fopen(tmpName.c_str(), "w");
    if (!fp) return __LINE__;
    std::fprintf(fp, "toa_s,freq_hz,pw_s,snr_db,sat,amp,channel\n");
    for (const chz_pdw_t& p : pdws)
      std::fprintf(fp, "%.9f,%.3f,%.9g,%.4f,%u,%.6g,%u\n", p.toa_s, p.freq_hz, p.pw_s, p.snr_db, p.saturated, p.amp,
                   p.channel);
    std::fclose(fp);
    if (std::rename(tmpName.c_str(), pdwName.c_str()) != 0) return __LINE__;
    std::cout << "Wrote " << chanName << " and " << pdwName << std::endl;
  }
  return 0;
}

bool isDirectory(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

bool exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}

std::vector<std::string> listIq(const std::string& dir) {
  std::vector<std::string> names;
  if (DIR* d = opendir(dir.c_str())) {
    while (const dirent* e = readdir(d)) {
      const std::string n = e->d_name;
      if (n.size() > 3 && n.compare(n.size() - 3, 3, ".iq") == 0) names.push_back(n);
    }
    closedir(d);
  }
  std::sort(names.begin(), names.end());
  return names;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 5) {
    std::cerr << "Usage: " << argv[0]
              << " <recording.iq | directory to watch> <channels (0 = sampleRate*1e-6)> <tapsPerBand> <oversample 1|2>"
                 " [snrThresholdDb=15] [outputPrefix | output directory] [idleSec=10 (watch mode: exit after this"
                 " long without a new file, or when a file named 'stop' appears)] [bugCompat=0 (1: pdw freq as"
                 " create_pdws_channelized.m:114 computes it, from column 1's phase)]"
              << std::endl;
    return __LINE__;
  }
  const std::string input = argv[1];
  Options opt;
  opt.channels = (uint32_t)std::atoi(argv[2]);
  opt.tapsPerBand = (uint32_t)std::atoi(argv[3]);
  opt.oversample = (uint32_t)std::atoi(argv[4]);
  opt.snrThresholdDb = argc > 5 ? std::atof(argv[5]) : 15.0;
  const std::string prefix = argc > 6 ? argv[6] : "";
  const double idleSec = argc > 7 ? std::atof(argv[7]) : 10.0;
  opt.bugCompat = argc > 8 && std::atoi(argv[8]) != 0;
  Engine eng;

  if (!isDirectory(input)) {
    int why = 0;
    const int rc = processOne(eng, opt, input, prefix, &why);
    if (rc == -1) {   // a truncated or unreadable recording is an error outside watch mode
      std::cerr << input << ": " << chz_strerror(why) << std::endl;
      return __LINE__;
    }
    return rc;
  }

  // watch mode: dwell files appear one by one while the recorder runs
  const std::string outDir = prefix.empty() ? input : prefix;
  std::set<std::string> done;
  // an incomplete file is waited for only while it keeps growing: a file that failed to open twice in a row
  // with the same size, at least a second apart, is reported and skipped so that it cannot block later dwells
  std::string stuckName;
  long long stuckSize = -1;
  auto stuckSince = std::chrono::steady_clock::now();
  auto lastWork = std::chrono::steady_clock::now();
  uint64_t processed = 0;
  for (;;) {
    bool worked = false;
    for (const std::string& name : listIq(input)) {
      if (done.count(name)) continue;
      const std::string stem = name.substr(0, name.size() - 3);
      int why = 0;
      const int rc = processOne(eng, opt, input + "/" + name, outDir + "/" + stem, &why);
      if (rc == -1) {                            // still being written: files are handled in time order, so wait for it
        struct stat st;
        const long long size = stat((input + "/" + name).c_str(), &st) == 0 ? (long long)st.st_size : -1;
        const auto now = std::chrono::steady_clock::now();
        if (name != stuckName || size != stuckSize) { stuckName = name; stuckSize = size; stuckSince = now; break; }
        if (std::chrono::duration<double>(now - stuckSince).count() < 1.0) break;
        std::cerr << input << "/" << name << ": skipped, " << chz_strerror(why) << " and the file stopped growing" << std::endl;
        done.insert(name);
        stuckName.clear();
        continue;
      }
      if (rc != 0) return rc;
      done.insert(name);
      processed++;
      worked = true;
    }
    const auto now = std::chrono::steady_clock::now();
    if (worked) lastWork = now;
    if (exists(input + "/stop")) break;
    if (std::chrono::duration<double>(now - lastWork).count() > idleSec) break;
    if (!worked) std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
  std::cout << "Processed " << processed << " recordings from " << input << std::endl;
  return 0;
}
