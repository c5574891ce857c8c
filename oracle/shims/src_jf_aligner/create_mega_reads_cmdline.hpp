// TEST INFRASTRUCTURE ONLY (oracle build). Hand-written stand-in for the
// header yaggo would generate from
// /root/reference/src_jf_aligner/create_mega_reads_cmdline.yaggo:1-84
// (field naming convention: <opt>_arg / <opt>_given / <flag>_flag, nested enum
// structs, static error() that prints and exit(1)s).
#ifndef ORACLE_SHIM_CREATE_MEGA_READS_CMDLINE_HPP
#define ORACLE_SHIM_CREATE_MEGA_READS_CMDLINE_HPP
#include "yaggo_shim.hpp"

class create_mega_reads_cmdline {
public:
  uint64_t size_arg;                  bool size_given;
  uint32_t mer_arg;                   bool mer_given;
  uint32_t fine_mer_arg;              bool fine_mer_given;
  uint32_t psa_min_arg;               bool psa_min_given;
  const char* unitigs_lengths_arg;    bool unitigs_lengths_given;
  const char* unitigs_sequences_arg;  bool unitigs_sequences_given;
  uint32_t k_mer_arg;                 bool k_mer_given;
  uint32_t threads_arg;               bool threads_given;
  const char* output_arg;             bool output_given;
  const char* dot_arg;                bool dot_given;
  int      stretch_constant_arg;      bool stretch_constant_given;
  double   stretch_factor_arg;        bool stretch_factor_given;
  double   stretch_cap_arg;           bool stretch_cap_given;
  uint32_t window_size_arg;           bool window_size_given;
  double   overlap_play_arg;          bool overlap_play_given;
  double   errors_arg;                bool errors_given;
  double   bases_matching_arg;        bool bases_matching_given;
  double   mers_matching_arg;         bool mers_matching_given;
  bool     max_match_flag;
  uint32_t max_count_arg;             bool max_count_given;
  bool     bases_flag;
  double   density_arg;               bool density_given;
  double   min_length_arg;            bool min_length_given;
  struct tiling { enum { none, greedy, maximal, weighted }; };
  int      tiling_arg;                bool tiling_given;
  struct trim { enum { none, match, branch }; };
  int      trim_arg;                  bool trim_given;
  std::vector<const char*> superreads_arg;
  std::vector<const char*> pacbio_arg;

  create_mega_reads_cmdline()
    : size_arg(0), size_given(false), mer_arg(0), mer_given(false), fine_mer_arg(0), fine_mer_given(false)
    , psa_min_arg(13), psa_min_given(false), unitigs_lengths_arg(""), unitigs_lengths_given(false)
    , unitigs_sequences_arg(""), unitigs_sequences_given(false), k_mer_arg(0), k_mer_given(false)
    , threads_arg(1), threads_given(false), output_arg(""), output_given(false), dot_arg(""), dot_given(false)
    , stretch_constant_arg(10), stretch_constant_given(false), stretch_factor_arg(1.3), stretch_factor_given(false)
    , stretch_cap_arg(10000.0), stretch_cap_given(false), window_size_arg(1), window_size_given(false)
    , overlap_play_arg(1.3), overlap_play_given(false), errors_arg(3.0), errors_given(false)
    , bases_matching_arg(17.0), bases_matching_given(false), mers_matching_arg(0.0), mers_matching_given(false)
    , max_match_flag(false), max_count_arg(5000), max_count_given(false), bases_flag(false)
    , density_arg(0.029), density_given(false), min_length_arg(100.0), min_length_given(false)
    , tiling_arg(tiling::greedy), tiling_given(false), trim_arg(trim::none), trim_given(false)
  { }

  static yaggo_shim::error_stream error() {
    return yaggo_shim::error_stream("Use --usage or --help for some help\n");
  }

  void parse(int argc, char* argv[]) {
    enum { O_PSA_MIN = 1000, O_DOT, O_SC, O_SF, O_SCAP, O_WS, O_MAXMATCH, O_MAXCOUNT, O_TRIM, O_USAGE };
    static struct option long_options[] = {
      {"size", 1, 0, 's'}, {"mer", 1, 0, 'm'}, {"fine-mer", 1, 0, 'F'}, {"psa-min", 1, 0, O_PSA_MIN},
      {"unitigs-lengths", 1, 0, 'l'}, {"unitigs-sequences", 1, 0, 'u'}, {"k-mer", 1, 0, 'k'},
      {"threads", 1, 0, 't'}, {"output", 1, 0, 'o'}, {"dot", 1, 0, O_DOT},
      {"stretch-constant", 1, 0, O_SC}, {"stretch-factor", 1, 0, O_SF}, {"stretch-cap", 1, 0, O_SCAP},
      {"window-size", 1, 0, O_WS}, {"overlap-play", 1, 0, 'O'}, {"errors", 1, 0, 'e'},
      {"bases-matching", 1, 0, 'B'}, {"mers-matching", 1, 0, 'M'}, {"max-match", 0, 0, O_MAXMATCH},
      {"max-count", 1, 0, O_MAXCOUNT}, {"bases", 0, 0, 'b'}, {"density", 1, 0, 'd'},
      {"min-length", 1, 0, 'L'}, {"tiling", 1, 0, 'T'}, {"trim", 1, 0, O_TRIM},
      {"superreads", 1, 0, 'r'}, {"pacbio", 1, 0, 'p'}, {"help", 0, 0, 'h'}, {"usage", 0, 0, O_USAGE},
      {"version", 0, 0, 'V'}, {0, 0, 0, 0}
    };
    static const char* short_options = "hVs:m:F:l:u:k:t:o:O:e:B:M:bd:L:T:r:p:";
    bool ok;
    optind = 1;
    while(true) {
      int index = -1;
      int c = getopt_long(argc, argv, short_options, long_options, &index);
      if(c == -1) break;
      switch(c) {
      case ':': case '?': error() << "Unrecognized or incomplete option"; break;
      case 'h': case O_USAGE: std::cout << "Usage: create_mega_reads [options]\n"; std::exit(0);
      case 'V': std::cout << "0.0.0\n"; std::exit(0);
      case 's': size_given = true; size_arg = yaggo_shim::conv_uint64(optarg, true, ok);
        if(!ok) error() << "Invalid uint64 '" << optarg << "' for [-s, --size=uint64]"; break;
      case 'm': mer_given = true; mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-m, --mer=uint32]"; break;
      case 'F': fine_mer_given = true; fine_mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-F, --fine-mer=uint32]"; break;
      case O_PSA_MIN: psa_min_given = true; psa_min_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [--psa-min=uint32]"; break;
      case 'l': unitigs_lengths_given = true; unitigs_lengths_arg = optarg; break;
      case 'u': unitigs_sequences_given = true; unitigs_sequences_arg = optarg; break;
      case 'k': k_mer_given = true; k_mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-k, --k-mer=uint32]"; break;
      case 't': threads_given = true; threads_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-t, --threads=uint32]"; break;
      case 'o': output_given = true; output_arg = optarg; break;
      case O_DOT: dot_given = true; dot_arg = optarg; break;
      case O_SC: stretch_constant_given = true; stretch_constant_arg = yaggo_shim::conv_int(optarg, ok);
        if(!ok) error() << "Invalid int '" << optarg << "' for [--stretch-constant=int]"; break;
      case O_SF: stretch_factor_given = true; stretch_factor_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [--stretch-factor=double]"; break;
      case O_SCAP: stretch_cap_given = true; stretch_cap_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [--stretch-cap=double]"; break;
      case O_WS: window_size_given = true; window_size_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [--window-size=uint32]"; break;
      case 'O': overlap_play_given = true; overlap_play_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-O, --overlap-play=double]"; break;
      case 'e': errors_given = true; errors_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-e, --errors=double]"; break;
      case 'B': bases_matching_given = true; bases_matching_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-B, --bases-matching=double]"; break;
      case 'M': mers_matching_given = true; mers_matching_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-M, --mers-matching=double]"; break;
      case O_MAXMATCH: max_match_flag = true; break;
      case O_MAXCOUNT: max_count_given = true; max_count_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [--max-count=uint32]"; break;
      case 'b': bases_flag = true; break;
      case 'd': density_given = true; density_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-d, --density=double]"; break;
      case 'L': min_length_given = true; min_length_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-L, --min-length=double]"; break;
      case 'T': tiling_given = true;
        if(!strcmp(optarg, "none")) tiling_arg = tiling::none;
        else if(!strcmp(optarg, "greedy")) tiling_arg = tiling::greedy;
        else if(!strcmp(optarg, "maximal")) tiling_arg = tiling::maximal;
        else if(!strcmp(optarg, "weighted")) tiling_arg = tiling::weighted;
        else error() << "Invalid enum '" << optarg << "' for [-T, --tiling]";
        break;
      case O_TRIM: trim_given = true;
        if(!strcmp(optarg, "none")) trim_arg = trim::none;
        else if(!strcmp(optarg, "match")) trim_arg = trim::match;
        else if(!strcmp(optarg, "branch")) trim_arg = trim::branch;
        else error() << "Invalid enum '" << optarg << "' for [--trim]";
        break;
      case 'r': superreads_arg.push_back(optarg); break;
      case 'p': pacbio_arg.push_back(optarg); break;
      }
    }
    if(!size_given) error() << "[-s, --size=uint64] required switch";
    if(!mer_given) error() << "[-m, --mer=uint32] required switch";
    if(!k_mer_given) error() << "[-k, --k-mer=uint32] required switch";
    if(unitigs_sequences_given && unitigs_lengths_given)
      error() << "Switches [-u, --unitigs-sequences=path] and [-l, --unitigs-lengths=path] are mutually exclusive";
    if(argc - optind != 0) error() << "Requires exactly 0 argument.";
  }
};
#endif
