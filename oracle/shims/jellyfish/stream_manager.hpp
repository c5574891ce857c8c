// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for
// <jellyfish/stream_manager.hpp>: hands out the input paths one by one
// (reference call sites: jf_aligner.hpp:24, create_mega_reads.cc:136).
#ifndef ORACLE_SHIM_JELLYFISH_STREAM_MANAGER_HPP
#define ORACLE_SHIM_JELLYFISH_STREAM_MANAGER_HPP
#include <fstream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
namespace jellyfish {
template<typename PathIterator>
class stream_manager {
  PathIterator cur_, end_;
  std::mutex   mutex_;
public:
  stream_manager(PathIterator b, PathIterator e) : cur_(b), end_(e) { }
  // next file, or null when exhausted
  std::unique_ptr<std::istream> next() {
    std::lock_guard<std::mutex> lock(mutex_);
    if(cur_ == end_) return std::unique_ptr<std::istream>();
    const char* path = *cur_;
    ++cur_;
    std::unique_ptr<std::istream> res(new std::ifstream(path));
    if(!res->good()) throw std::runtime_error(std::string("Can't open file '") + path + "'");
    return res;
  }
};
}
#endif
