// channelize_iq — command-line companion of the reference's record tools: takes a recording written
// by blade_record_iq_* / usrp_record_iq_* (cpp/IqPacket.h header + interleaved I/Q), channelizes it on
// the GPU through libchannelizer and writes the channel matrix and the PDW table.
// Style follows the reference tools: positional arguments, a usage text, `return __LINE__` on failure
// (cpp/blade_record_iq_12bit.cpp:31-37,54-59).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "channelizer.h"

#define CHECK(call)                                                                              \
  do {                                                                                           \
    const int status = (call);                                                                   \
    if (status != 0) {                                                                           \
      std::cerr << #call << " failed: " << chz_strerror(status) << " " << chz_last_cuda_error()  \
                << std::endl;                                                                    \
      return __LINE__;                                                                           \
    }                                                                                            \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 5) {
    std::cerr << "Usage: " << argv[0]
              << " <recording.iq> <channels (0 = sampleRate*1e-6)> <tapsPerBand> <oversample 1|2>"
                 " [snrThresholdDb=15] [outputPrefix]"
              << std::endl;
    return __LINE__;
  }
  const char* path = argv[1];
  uint32_t channels = (uint32_t)std::atoi(argv[2]);
  const uint32_t tapsPerBand = (uint32_t)std::atoi(argv[3]);
  const uint32_t oversample = (uint32_t)std::atoi(argv[4]);
  const double snrThresholdDb = argc > 5 ? std::atof(argv[5]) : 15.0;
  const std::string prefix = argc > 6 ? argv[6] : "";

  chz_iq_t* file = nullptr;
  chz_iq_info_t info;
  CHECK(chz_open_iq(path, &file, &info));
  std::cout << "File format " << info.format << ", " << info.num_samples << " samples, " << info.bit_width
            << " bits, fs " << info.fs_sps << " sps, fc " << info.fc_hz << " Hz, board '" << info.board_name << "'"
            << std::endl;
  if (channels == 0) channels = (uint32_t)(info.fs_sps * 1e-6 + 0.5);   // create_pdws_channelized.m:31

  std::vector<float> taps((size_t)channels * tapsPerBand);
  CHECK(chz_design_prototype(channels, tapsPerBand, 80.0, taps.data()));
  chz_t* chan = nullptr;
  CHECK(chz_create(channels, taps.data(), (uint32_t)taps.size(), oversample, &chan));

  const uint64_t rows = chz_rows_for(chan, info.num_samples);
  chz_cf32* out = nullptr;
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
  prm.snr_threshold_db = snrThresholdDb;
  prm.sat_level = 0.9999;
  prm.fc_hz = (double)info.fc_hz;
  prm.fs_sps = (double)info.fs_sps;
  prm.t0 = info.sample_start_time;
  prm.use_trailing_threshold = 0;
  uint64_t npdw = 0;
  int status = chz_pdws(chan, &prm, nullptr, 0, &npdw);
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
    const std::string pdwName = prefix + ".pdw.csv";
    FILE* fp = std::fopen(pdwName.c_str(), "w");
    if (!fp) return __LINE__;
    std::fprintf(fp, "toa_s,freq_hz,pw_s,snr_db,sat,amp,channel\n");
    for (const chz_pdw_t& p : pdws)
      std::fprintf(fp, "%.9f,%.3f,%.9g,%.4f,%u,%.6g,%u\n", p.toa_s, p.freq_hz, p.pw_s, p.snr_db, p.saturated, p.amp,
                   p.channel);
    std::fclose(fp);
    std::cout << "Wrote " << chanName << " and " << pdwName << std::endl;
  }
  if (out) chz_free_host(out);
  chz_destroy(chan);
  chz_close_iq(file);
  return 0;
}
