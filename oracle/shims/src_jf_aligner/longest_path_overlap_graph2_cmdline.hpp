// TEST INFRASTRUCTURE ONLY (oracle build). Hand-written stand-in for the header yaggo would
// generate from /root/reference/src_jf_aligner/longest_path_overlap_graph2_cmdline.yaggo:1-49.
#ifndef ORACLE_SHIM_LONGEST_PATH_CMDLINE_HPP
#define ORACLE_SHIM_LONGEST_PATH_CMDLINE_HPP
#include "yaggo_shim.hpp"

class longest_path_overlap_graph2_cmdline {
public:
  struct tiling { enum { none, greedy, maximal }; };
  struct trim   { enum { none, match, branch }; };
  uint32_t threads_arg;                bool threads_given;
  const char* output_arg;              bool output_given;
  const char* dot_arg;                 bool dot_given;
  double   overlap_play_arg;           bool overlap_play_given;
  double   errors_arg;                 bool errors_given;
  bool     bases_flag;
  const char* unitigs_lengths_arg;     bool unitigs_lengths_given;
  const char* unitigs_sequences_arg;   bool unitigs_sequences_given;
  uint32_t k_mer_arg;                  bool k_mer_given;
  double   density_arg;                bool density_given;
  double   min_length_arg;             bool min_length_given;
  int      tiling_arg;                 bool tiling_given;
  int      trim_arg;                   bool trim_given;
  const char* coords_arg;

  longest_path_overlap_graph2_cmdline()
    : threads_arg(1), threads_given(false), output_arg(""), output_given(false), dot_arg(""), dot_given(false)
    , overlap_play_arg(1.3), overlap_play_given(false), errors_arg(3.0), errors_given(false), bases_flag(false)
    , unitigs_lengths_arg(""), unitigs_lengths_given(false), unitigs_sequences_arg(""), unitigs_sequences_given(false)
    , k_mer_arg(0), k_mer_given(false), density_arg(0.029), density_given(false), min_length_arg(100.0), min_length_given(false)
    , tiling_arg(tiling::greedy), tiling_given(false), trim_arg(trim::none), trim_given(false), coords_arg("")
  { }

  static yaggo_shim::error_stream error() {
    return yaggo_shim::error_stream("Use --usage or --help for some help\n");
  }

  void parse(int argc, char* argv[]) {
    enum { O_DOT = 1000, O_TRIM, O_USAGE };
    static struct option long_options[] = {
      {"threads", 1, 0, 't'}, {"output", 1, 0, 'o'}, {"dot", 1, 0, O_DOT}, {"overlap-play", 1, 0, 'O'},
      {"errors", 1, 0, 'e'}, {"bases", 0, 0, 'b'}, {"unitigs-lengths", 1, 0, 'l'}, {"unitigs-sequences", 1, 0, 'u'},
      {"k-mer", 1, 0, 'k'}, {"density", 1, 0, 'd'}, {"min-length", 1, 0, 'L'}, {"tiling", 1, 0, 'T'},
      {"trim", 1, 0, O_TRIM}, {"help", 0, 0, 'h'}, {"usage", 0, 0, O_USAGE}, {"version", 0, 0, 'V'}, {0, 0, 0, 0}
    };
    bool ok;
    optind = 1;
    while(true) {
      int index = -1;
      int c = getopt_long(argc, argv, "hVt:o:O:e:bl:u:k:d:L:T:", long_options, &index);
      if(c == -1) break;
      switch(c) {
      case ':': case '?': error() << "Unrecognized or incomplete option"; break;
      case 'h': case O_USAGE: std::cout << "Usage: longest_path_overlap_graph2 [options] coords:path\n"; std::exit(0);
      case 'V': std::cout << "0.0.0\n"; std::exit(0);
      case 't': threads_given = true; threads_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-t, --threads=uint32]"; break;
      case 'o': output_given = true; output_arg = optarg; break;
      case O_DOT: dot_given = true; dot_arg = optarg; break;
      case 'O': overlap_play_given = true; overlap_play_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-O, --overlap-play=double]"; break;
      case 'e': errors_given = true; errors_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-e, --errors=double]"; break;
      case 'b': bases_flag = true; break;
      case 'l': unitigs_lengths_given = true; unitigs_lengths_arg = optarg; break;
      case 'u': unitigs_sequences_given = true; unitigs_sequences_arg = optarg; break;
      case 'k': k_mer_given = true; k_mer_arg = yaggo_shim::conv_uint64(optarg, false, ok);
        if(!ok) error() << "Invalid uint32 '" << optarg << "' for [-k, --k-mer=uint32]"; break;
      case 'd': density_given = true; density_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-d, --density=double]"; break;
      case 'L': min_length_given = true; min_length_arg = yaggo_shim::conv_double(optarg, ok);
        if(!ok) error() << "Invalid double '" << optarg << "' for [-L, --min-length=double]"; break;
      case 'T': tiling_given = true;
        if(!strcmp(optarg, "none")) tiling_arg = tiling::none; else if(!strcmp(optarg, "greedy")) tiling_arg = tiling::greedy;
        else if(!strcmp(optarg, "maximal")) tiling_arg = tiling::maximal;
        else error() << "Invalid enum '" << optarg << "' for [-T, --tiling=enum]"; break;
      case O_TRIM: trim_given = true;
        if(!strcmp(optarg, "none")) trim_arg = trim::none; else if(!strcmp(optarg, "match")) trim_arg = trim::match;
        else if(!strcmp(optarg, "branch")) trim_arg = trim::branch;
        else error() << "Invalid enum '" << optarg << "' for [--trim=enum]"; break;
      }
    }
    if(!k_mer_given) error() << "[-k, --k-mer=uint32] required switch";
    if(unitigs_sequences_given && unitigs_lengths_given)
      error() << "Switches [-u, --unitigs-sequences=path] and [-l, --unitigs-lengths=path] are mutually exclusive";
    if(argc - optind != 1) error() << "Requires exactly 1 argument.";
    coords_arg = argv[optind];
  }
};
#endif
