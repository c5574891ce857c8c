// TEST INFRASTRUCTURE ONLY (oracle build). Hand-written stand-in for the
// header yaggo would generate from
// /root/reference/src_jf_aligner/jf_aligner_cmdline.yaggo:1-77.
#ifndef ORACLE_SHIM_JF_ALIGNER_CMDLINE_HPP
#define ORACLE_SHIM_JF_ALIGNER_CMDLINE_HPP
#include "yaggo_shim.hpp"

class jf_aligner_cmdline {
public:
  uint64_t size_arg;                  bool size_given;
  uint32_t mer_arg;                   bool mer_given;
  uint32_t fine_mer_arg;              bool fine_mer_given;
  uint32_t psa_min_arg;               bool psa_min_given;
  uint32_t threads_arg;               bool threads_given;
  int      stretch_constant_arg;      bool stretch_constant_given;
  double   stretch_factor_arg;        bool stretch_factor_given;
  double   stretch_cap_arg;           bool stretch_cap_given;
  uint32_t window_size_arg;           bool window_size_given;
  bool     forward_flag;
  double   bases_matching_arg;        bool bases_matching_given;
  double   mers_matching_arg;         bool mers_matching_given;
  const char* details_arg;            bool details_given;
  const char* coords_arg;             bool coords_given;
  bool     max_match_flag;
  bool     no_header_flag;
  bool     zero_match_flag;
  uint32_t max_count_arg;             bool max_count_given;
  const char* unitigs_lengths_arg;    bool unitigs_lengths_given;
  const char* unitigs_sequences_arg;  bool unitigs_sequences_given;
  bool     compact_flag;
  uint32_t k_mer_arg;                 bool k_mer_given;
  std::vector<const char*> superreads_arg;
  std::vector<const char*> pacbio_arg;

  jf_aligner_cmdline()
    : size_arg(0), size_given(false), mer_arg(0), mer_given(false), fine_mer_arg(0), fine_mer_given(false)
    , psa_min_arg(13), psa_min_given(false), threads_arg(1), threads_given(false)
    , stretch_constant_arg(10), stretch_constant_given(false), stretch_factor_arg(1.3), stretch_factor_given(false)
    , stretch_cap_arg(10000.0), stretch_cap_given(false), window_size_arg(1), window_size_given(false)
    , forward_flag(false), bases_matching_arg(17.0), bases_matching_given(false)
    , mers_matching_arg(0.0), mers_matching_given(false), details_arg(""), details_given(false)
    , coords_arg(""), coords_given(false), max_match_flag(false), no_header_flag(false), zero_match_flag(false)
    , max_count_arg(5000), max_count_given(false), unitigs_lengths_arg(""), unitigs_lengths_given(false)
    , unitigs_sequences_arg(""), unitigs_sequences_given(false), compact_flag(true)
    , k_mer_arg(0), k_mer_given(false)
  { }

  static yaggo_shim::error_stream error() {
    return yaggo_shim::error_stream("Use --usage or --help for some help\n");
  }

  void parse(int argc, char* argv[]) {
    enum { O_PSA_MIN = 1000, O_SC, O_SF, O_SCAP, O_WS, O_DETAILS, O_COORDS, O_MAXMATCH, O_MAXCOUNT, O_COMPACT, O_USAGE };
    static struct option long_options[] = {
      {"size", 1, 0, 's'}, {"mer", 1, 0, 'm'}, {"fine-mer", 1, 0, 'F'}, {"psa-min", 1, 0, O_PSA_MIN},
      {"threads", 1, 0, 't'}, {"stretch-constant", 1, 0, O_SC}, {"stretch-factor", 1, 0, O_SF},
      {"stretch-cap", 1, 0, O_SCAP}, {"window-size", 1, 0, O_WS}, {"forward", 0, 0, 'f'},
      {"bases-matching", 1, 0, 'B'}, {"mers-matching", 1, 0, 'M'}, {"details", 1, 0, O_DETAILS},
      {"coords", 1, 0, O_COORDS}, {"max-match", 0, 0, O_MAXMATCH}, {"no-header", 0, 0, 'H'},
      {"zero-match", 0, 0, '0'}, {"max-count", 1, 0, O_MAXCOUNT}, {"unitigs-lengths", 1, 0, 'l'},
      {"unitigs-sequences", 1, 0, 'u'}, {"compact", 0, 0, O_COMPACT}, {"k-mer", 1, 0, 'k'},
      {"superreads", 1, 0, 'r'}, {"pacbio", 1, 0, 'p'}, {"help", 0, 0, 'h'}, {"usage", 0, 0, O_USAGE},
      {"version", 0, 0, 'V'}, {0, 0, 0, 0}
    };
    static const char* short_options = "hVs:m:F:t:fB:M:H0l:u:k:r:p:";
    bool ok;
    optind = 1;
    while(true) {
      int index = -1;
      int c = getopt_long(argc, argv, short_options, long_options, &index);
      if(c == -1) break;
      switch(c) {
      case ':': case '?': error() << "Unrecognized or incomplete option"; break;
      case 'h': case O_USAGE: std::cout << "Usage: jf_aligner [options]\n"; std::exit(0);
      case 'V': std::cout << "0.0.0\n"; std::exit(0);
      case 's': size_given = true; size_arg = yaggo_shim::conv_uint64(optarg, true, ok);
        if(!ok) error() << "Invalid uint64 '" << optarg << "' for [-s, --size=uint64]"; break;
      case 'm': mer_given = true; mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-m, --mer=uint32]"; break;
      case 'F': fine_mer_given = true; fine_mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-F, --fine-mer=uint32]"; break;
      case O_PSA_MIN: psa_min_given = true; psa_min_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [--psa-min=uint32]"; break;
      case 't': threads_given = true; threads_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-t, --threads=uint32]"; break;
      case O_SC: stretch_constant_given = true; stretch_constant_arg = yaggo_shim::conv_int(optarg, ok);
        if(!ok) error() << "Invalid int '" << optarg << "' for [--stretch-constant=int]"; break;
      case O_SF: stretch_factor_given = true; stretch_factor_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [--stretch-factor=double]"; break;
      case O_SCAP: stretch_cap_given = true; stretch_cap_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [--stretch-cap=double]"; break;
      case O_WS: window_size_given = true; window_size_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [--window-size=uint32]"; break;
      case 'f': forward_flag = true; break;
      case 'B': bases_matching_given = true; bases_matching_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-B, --bases-matching=double]"; break;
      case 'M': mers_matching_given = true; mers_matching_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-M, --mers-matching=double]"; break;
      case O_DETAILS: details_given = true; details_arg = optarg; break;
      case O_COORDS: coords_given = true; coords_arg = optarg; break;
      case O_MAXMATCH: max_match_flag = true; break;
      case 'H': no_header_flag = true; break;
      case '0': zero_match_flag = true; break;
      case O_MAXCOUNT: max_count_given = true; max_count_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [--max-count=uint32]"; break;
      case 'l': unitigs_lengths_given = true; unitigs_lengths_arg = optarg; forward_flag = true; break;
      case 'u': unitigs_sequences_given = true; unitigs_sequences_arg = optarg; forward_flag = true; break;
      case O_COMPACT: compact_flag = false; break;
      case 'k': k_mer_given = true; k_mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-k, --k-mer=uint32]"; break;
      case 'r': superreads_arg.push_back(optarg); break;
      case 'p': pacbio_arg.push_back(optarg); break;
      }
    }
    if(!size_given) error() << "[-s, --size=uint64] required switch";
    if(!mer_given) error() << "[-m, --mer=uint32] required switch";
    if(unitigs_sequences_given && unitigs_lengths_given)
      error() << "Switches [-u, --unitigs-sequences=path] and [-l, --unitigs-lengths=path] are mutually exclusive";
    if(argc - optind != 0) error() << "Requires exactly 0 argument.";
  }
};
#endif
