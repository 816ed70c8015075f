// Writes a tiny recording using the REFERENCE's own header struct (IqPacket, compiled from
// /root/reference/cpp/IqPacket.h — included, never copied) exactly the way the recorders do:
// fout.write(&packet, sizeof(packet)); fout.write(iq, n*sizeof(complex<int16>))
// (cpp/blade_record_iq_12bit.cpp:320-323, cpp/usrp_record_iq_12bit.cpp:224-227).
// Built only where /root/reference exists (oracle/Makefile target `ref`); the output is committed
// as tests/golden/ref_iqpacket_fmt3.iq and pins the header layout the reader must parse.
// Also exercises the reference's Helper.cpp getFilenameStr() for the CLI's naming test.
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "IqPacket.h"
#include "Helper.h"

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: " << argv[0] << " <out.iq> [bitWidth=12]\n"; return __LINE__; }
  const std::uint32_t bitWidth = argc > 2 ? std::atoi(argv[2]) : 12;
  IqPacket packet;
  std::memset(&packet, 0, sizeof(packet));
  packet.endianness = 0x03030303;          // what the current recorders write for IQ_FILE_FORMAT 3
  packet.linkSpeed = 5000;
  packet.frequencyHz = 5800000000ULL;      // > 2^32 on purpose (format 1 could not hold it)
  packet.bandwidthHz = 56000000;
  packet.sampleRateSps = 61440000;
  packet.rxGainDb = 37.5f;
  packet.bitWidth = bitWidth;
  packet.spare0 = 0;
  std::strncpy(packet.boardName, "bladerf2", sizeof(packet.boardName));
  std::strncpy(packet.serialNumber, "0123456789abcdef", sizeof(packet.serialNumber));  // full 16, no NUL
  std::strncpy(packet.fpgaVersion, "0.15.0", sizeof(packet.fpgaVersion));
  std::strncpy(packet.fwVersion, "2.4.0", sizeof(packet.fwVersion));
  packet.sampleStartTime = 1700000000.123456;
  const std::uint32_t n = 37;
  packet.numSamples = n;
  std::ofstream fout(argv[1], std::ios::binary);
  fout.write((const char*)&packet, sizeof(packet));
  if (bitWidth <= 8) {
    std::complex<std::int8_t> iq[n];
    for (std::uint32_t i = 0; i < n; i++) iq[i] = std::complex<std::int8_t>((std::int8_t)(i * 7 - 128), (std::int8_t)(127 - i * 5));
    fout.write((const char*)iq, n * sizeof(iq[0]));
  } else {
    std::complex<std::int16_t> iq[n];
    for (std::uint32_t i = 0; i < n; i++) iq[i] = std::complex<std::int16_t>((std::int16_t)(i * 113 - 2048), (std::int16_t)(2047 - i * 97));
    fout.write((const char*)iq, n * sizeof(iq[0]));
  }
  fout.close();
  char name[64];
  getFilenameStr(std::chrono::system_clock::from_time_t(1700000000) + std::chrono::milliseconds(123), name, sizeof name);
  std::cout << "sizeof(IqPacket)=" << sizeof(IqPacket) << " filename=" << name << "\n";
  return 0;
}
