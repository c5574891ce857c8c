// TEST INFRASTRUCTURE ONLY (oracle build). Hand-written stand-in for the header yaggo would
// generate from /root/reference/src_jf_aligner/merge_coords_cmdline.yaggo:1-15.
#ifndef ORACLE_SHIM_MERGE_COORDS_CMDLINE_HPP
#define ORACLE_SHIM_MERGE_COORDS_CMDLINE_HPP
#include "yaggo_shim.hpp"
class merge_coords_cmdline {
public:
  const char* output_arg; bool output_given;
  std::vector<const char*> coords_arg;
  merge_coords_cmdline(int argc, char* argv[]) : output_arg(""), output_given(false) { parse(argc, argv); }
  static yaggo_shim::error_stream error() { return yaggo_shim::error_stream("Use --usage or --help for some help\n"); }
  void parse(int argc, char* argv[]) {
    static struct option long_options[] = { {"output", 1, 0, 'o'}, {"help", 0, 0, 'h'}, {"usage", 0, 0, 'U'}, {"version", 0, 0, 'V'}, {0, 0, 0, 0} };
    optind = 1;
    while(true) {
      int c = getopt_long(argc, argv, "hVo:", long_options, 0);
      if(c == -1) break;
      switch(c) {
      case ':': case '?': error() << "Unrecognized or incomplete option"; break;
      case 'h': case 'U': std::cout << "Usage: merge_coords [options] coords:PATH+\n"; std::exit(0);
      case 'V': std::cout << "0.0.0\n"; std::exit(0);
      case 'o': output_given = true; output_arg = optarg; break;
      }
    }
    for(int i = optind; i < argc; ++i) coords_arg.push_back(argv[i]);
  }
};
#endif
