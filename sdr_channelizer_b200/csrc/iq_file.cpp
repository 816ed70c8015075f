// R1: reader/writer for the reference's .iq recording format.
// Reader semantics follow matlab/convert_my_iq_to_mat.m:38-102 (the reference's only parser); the
// byte layout of formats 2/3 is the IqPacket struct the recorders dump (cpp/IqPacket.h:9-25,
// cpp/blade_record_iq_12bit.cpp:320-323); format 1 is what matlab/generate_training_iq.m:107-125
// writes (u32 centre frequency, no spare0 => 104-byte header).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "channelizer.h"

struct chz_iq {
  int fd;
  void* map;
  size_t map_bytes;
  chz_iq_info_t info;
};

namespace {

// Little-endian field readers (the reference only ever produces native LE files; the "big endian"
// magic 0 is accepted but never byte-swapped, convert_my_iq_to_mat.m:43-45).
struct Cursor {
  const unsigned char* p;
  size_t off;
  uint32_t u32() { uint32_t v; memcpy(&v, p + off, 4); off += 4; return v; }
  uint64_t u64() { uint64_t v; memcpy(&v, p + off, 8); off += 8; return v; }
  float f32() { float v; memcpy(&v, p + off, 4); off += 4; return v; }
  double f64() { double v; memcpy(&v, p + off, 8); off += 8; return v; }
  void str16(char* dst) {   // strip(string(fread(fid,16,'*char')'),char(0)): drop leading/trailing NULs
    const unsigned char* s = p + off;
    int b = 0, e = 16;
    while (b < e && s[b] == 0) b++;
    while (e > b && s[e - 1] == 0) e--;
    memcpy(dst, s + b, (size_t)(e - b));
    dst[e - b] = 0;
    off += 16;
  }
};

int parse_header(const unsigned char* bytes, uint64_t file_bytes, chz_iq_info_t* o) {
  memset(o, 0, sizeof *o);
  if (file_bytes < 4) return CHZ_EIO;
  Cursor c{bytes, 0};
  o->magic = c.u32();
  if (o->magic == 0x01010101u) o->format = 1;
  else if (o->magic == 0x02020202u || o->magic == 0u) o->format = 2;
  else if (o->magic == 0x03030303u) o->format = 3;
  else return CHZ_EFORMAT;
  o->header_bytes = o->format == 1 ? 104u : 112u;
  if (file_bytes < o->header_bytes) return CHZ_EIO;
  o->link_speed = c.u32();
  o->fc_hz = o->format == 1 ? (uint64_t)c.u32() : c.u64();
  o->bw_hz = c.u32();
  o->fs_sps = c.u32();
  o->gain_db = o->format >= 3 ? (double)c.f32() : (double)c.u32();
  o->num_samples = c.u32();
  o->bit_width = c.u32();
  o->spare0 = o->format >= 2 ? c.u32() : 0u;
  c.str16(o->board_name);
  c.str16(o->serial_number);
  c.str16(o->fpga_version);
  c.str16(o->fw_version);
  o->sample_start_time = c.f64();
  if (o->bit_width > 0 && o->bit_width <= 8) o->bytes_per_sample = 2;
  else if (o->bit_width > 8 && o->bit_width <= 16) o->bytes_per_sample = 4;
  else return CHZ_EBITWIDTH;
  o->payload_offset = c.off;
  const uint64_t pairs = (file_bytes - c.off) / o->bytes_per_sample;   // fread(fid,[2,inf]) keeps whole pairs
  if (pairs != o->num_samples) return CHZ_ESIZE;                       // assert(length(iq) == numSamples)
  o->payload_bytes = pairs * o->bytes_per_sample;
  return CHZ_OK;
}

}  // namespace

extern "C" int chz_open_iq(const char* path, chz_iq_t** out, chz_iq_info_t* info) {
  if (!path || (!out && !info)) return CHZ_EINVAL;
  if (out) *out = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return CHZ_EIO;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return CHZ_EIO; }
  void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (map == MAP_FAILED) { close(fd); return CHZ_EIO; }
  chz_iq_info_t tmp;
  const int rc = parse_header((const unsigned char*)map, (uint64_t)st.st_size, &tmp);
  if (info) *info = tmp;
  if (rc != CHZ_OK || !out) {
    munmap(map, (size_t)st.st_size);
    close(fd);
    return rc;
  }
  chz_iq* f = new (std::nothrow) chz_iq;
  if (!f) { munmap(map, (size_t)st.st_size); close(fd); return CHZ_ENOMEM; }
  f->fd = fd; f->map = map; f->map_bytes = (size_t)st.st_size; f->info = tmp;
  madvise(map, f->map_bytes, MADV_SEQUENTIAL);
  *out = f;
  return CHZ_OK;
}

extern "C" const void* chz_iq_payload(const chz_iq_t* f) {
  return f ? (const unsigned char*)f->map + f->info.payload_offset : nullptr;
}

extern "C" int chz_close_iq(chz_iq_t* f) {
  if (!f) return CHZ_EINVAL;
  munmap(f->map, f->map_bytes);
  close(f->fd);
  delete f;
  return CHZ_OK;
}

extern "C" int chz_write_iq(const char* path, const chz_iq_info_t* in, const void* payload) {
  if (!path || !in || (!payload && in->num_samples)) return CHZ_EINVAL;
  if (in->format < 1 || in->format > 3) return CHZ_EINVAL;
  if (in->bit_width == 0 || in->bit_width > 16) return CHZ_EBITWIDTH;
  unsigned char hdr[112];
  memset(hdr, 0, sizeof hdr);
  size_t off = 0;
  auto put = [&](const void* v, size_t n) { memcpy(hdr + off, v, n); off += n; };
  const uint32_t magic = in->format == 1 ? 0x01010101u : in->format == 2 ? 0x02020202u : 0x03030303u;
  put(&magic, 4);
  put(&in->link_speed, 4);
  if (in->format == 1) { const uint32_t fc = (uint32_t)in->fc_hz; put(&fc, 4); } else put(&in->fc_hz, 8);
  put(&in->bw_hz, 4);
  put(&in->fs_sps, 4);
  if (in->format >= 3) { const float g = (float)in->gain_db; put(&g, 4); } else { const uint32_t g = (uint32_t)in->gain_db; put(&g, 4); }
  put(&in->num_samples, 4);
  put(&in->bit_width, 4);
  if (in->format >= 2) put(&in->spare0, 4);
  const char* strs[4] = {in->board_name, in->serial_number, in->fpga_version, in->fw_version};
  for (const char* s : strs) { char b[16]; memset(b, 0, 16); memcpy(b, s, strnlen(s, 16)); put(b, 16); }
  put(&in->sample_start_time, 8);
  FILE* fp = fopen(path, "wb");
  if (!fp) return CHZ_EIO;
  const size_t bps = in->bit_width <= 8 ? 2 : 4;
  bool ok = fwrite(hdr, 1, off, fp) == off;
  if (ok && in->num_samples) ok = fwrite(payload, bps, in->num_samples, fp) == in->num_samples;
  ok = (fclose(fp) == 0) && ok;
  return ok ? CHZ_OK : CHZ_EIO;
}
