// Minimal stand-in for the option structs yaggo generates in the reference build
// (create_mega_reads_cmdline.yaggo, jf_aligner_cmdline.yaggo): long/short options, k/M/G suffix
// on uint64, error() that prints the message plus a usage hint and exits 1.
#pragma once
#include <getopt.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace cmdline {

[[noreturn]] inline void error(const std::string& msg) {
  fprintf(stderr, "Error: %s\nUse --usage or --help for some help\n\n", msg.c_str());
  exit(1);
}
inline uint64_t to_uint64(const char* s, const char* what, bool suffix) {
  char* end = nullptr;
  if(!s || !*s || *s == '-') error(std::string("Invalid uint64 '") + (s ? s : "") + "' for [" + what + "]");
  unsigned long long v = strtoull(s, &end, 0);
  if(end == s) error(std::string("Invalid uint64 '") + s + "' for [" + what + "]");
  if(*end && suffix) {
    switch(*end) {
    case 'k': v *= 1000ULL; ++end; break;
    case 'M': v *= 1000000ULL; ++end; break;
    case 'G': v *= 1000000000ULL; ++end; break;
    case 'T': v *= 1000000000000ULL; ++end; break;
    default: break;
    }
  }
  if(*end) error(std::string("Invalid uint64 '") + s + "' for [" + what + "]");
  return v;
}
inline uint32_t to_uint32(const char* s, const char* what) { return (uint32_t)to_uint64(s, what, false); }
inline double to_double(const char* s, const char* what) {
  char* end = nullptr;
  const double v = strtod(s, &end);
  if(end == s || *end) error(std::string("Invalid double '") + s + "' for [" + what + "]");
  return v;
}
inline long to_int(const char* s, const char* what) {
  char* end = nullptr;
  const long v = strtol(s, &end, 0);
  if(end == s || *end) error(std::string("Invalid int '") + s + "' for [" + what + "]");
  return v;
}

} // namespace cmdline
