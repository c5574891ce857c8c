// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for
// <jellyfish/whole_sequence_parser.hpp>: FASTA/FASTQ records in groups of
// `group` reads handed to worker threads as jobs (reference call sites:
// jf_aligner.hpp:25, create_mega_reads.cc:51-58,137).  Header excludes the
// leading '>'/'@'; sequence has its line breaks removed.
#ifndef ORACLE_SHIM_JELLYFISH_WHOLE_SEQUENCE_PARSER_HPP
#define ORACLE_SHIM_JELLYFISH_WHOLE_SEQUENCE_PARSER_HPP
#include <istream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <stdexcept>
namespace jellyfish {
struct header_sequence_qual {
  std::string header, seq, qual;
};
struct sequence_list {
  size_t nb_filled;
  std::vector<header_sequence_qual> data;
};

template<typename StreamManager>
class whole_sequence_parser {
  StreamManager&                 streams_;
  std::unique_ptr<std::istream>  cur_;
  std::mutex                     mutex_;
  const size_t                   group_;

  bool fill_one(header_sequence_qual& r) {
    while(true) {
      if(!cur_) { cur_ = streams_.next(); if(!cur_) return false; }
      int c = cur_->peek();
      while(c == '\n' || c == '\r') { cur_->get(); c = cur_->peek(); }
      if(c == EOF) { cur_.reset(); continue; }
      if(c == '>') {
        cur_->get();
        std::getline(*cur_, r.header);
        r.seq.clear(); r.qual.clear();
        std::string line;
        for(c = cur_->peek(); c != '>' && c != EOF; c = cur_->peek()) {
          std::getline(*cur_, line);
          r.seq += line;
        }
        return true;
      } else if(c == '@') {
        cur_->get();
        std::getline(*cur_, r.header);
        r.seq.clear(); r.qual.clear();
        std::string line;
        for(c = cur_->peek(); c != '+' && c != EOF; c = cur_->peek()) {
          std::getline(*cur_, line);
          r.seq += line;
        }
        if(c == '+') {
          std::getline(*cur_, line);
          while(r.qual.size() < r.seq.size() && cur_->good()) {
            std::getline(*cur_, line);
            r.qual += line;
          }
        }
        return true;
      } else {
        throw std::runtime_error("Unsupported format");
      }
    }
  }

public:
  whole_sequence_parser(uint32_t /*queue*/, uint32_t group, uint32_t /*max_producers*/, StreamManager& streams)
    : streams_(streams), group_(group) { }

  class job {
    std::unique_ptr<sequence_list> list_;
  public:
    explicit job(whole_sequence_parser& p) : list_(new sequence_list) {
      list_->nb_filled = 0;
      list_->data.resize(p.group_);
      std::lock_guard<std::mutex> lock(p.mutex_);
      while(list_->nb_filled < p.group_ && p.fill_one(list_->data[list_->nb_filled]))
        ++list_->nb_filled;
    }
    bool is_empty() const { return list_->nb_filled == 0; }
    sequence_list* operator->() { return list_.get(); }
    sequence_list& operator*() { return *list_; }
  };
  friend class job;
};
}
#endif
