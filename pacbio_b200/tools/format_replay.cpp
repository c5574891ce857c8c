// Times the host formatting stage (mega_reads_per_comp + tiling + printing, host_common.cpp) on a result
// dumped by MR_DUMP_BATCH=<file> -- no GPU needed, so the stage can be profiled and tuned on any host.
//   format_replay <dump> <superreads.fa> <unitigs.fa> <unitig_k> <threads> [reps] [out.txt]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../csrc/host/host_common.hpp"

int main(int argc, char** argv) {
  if(argc < 6) { fprintf(stderr, "usage: %s dump superreads.fa unitigs.fa unitig_k threads [reps] [out]\n", argv[0]); return 1; }
  try {
    mrh::result_dump d;
    if(!mrh::load_result(argv[1], d)) throw std::runtime_error("cannot read the dump");
    mrh::super_reads SR;
    SR.append_fasta(argv[2]);
    mrh::unitigs U;
    U.load_sequences(argv[3]);
    mrh::graph_options G;
    G.k_len = (uint32_t)atoi(argv[4]);
    if(const char* e = getenv("MR_REPLAY_TILING")) G.tiling = atoi(e);       // 0 none, 1 greedy, 2 maximal, 3 weighted
    if(const char* e = getenv("MR_REPLAY_TRIM")) G.trim = atoi(e);
    const unsigned threads = (unsigned)atoi(argv[5]);
    const int reps = argc > 6 ? atoi(argv[6]) : 5;
    std::vector<mrh::text_buf> parts;
    double best = 1e30;
    uint64_t bytes = 0;
    for(int i = 0; i < reps; ++i) {
      const auto t0 = std::chrono::steady_clock::now();
      // MR_REPLAY_SLICES=1: slice by slice through an emit callback, as the tools and bench.py's end-to-end arm do
      std::string sliced;
      uint64_t emitted = 0;
      const mrh::emit_fn emit = [&](std::vector<mrh::text_buf>& ps) { for(auto& p : ps) { emitted += p.size(); if(argc > 7 && i == 0) sliced.append(p.data(), p.size()); } };
      const bool slices = getenv("MR_REPLAY_SLICES") != nullptr;
      mrh::format_mega_reads_mt(d.view, d.batch, SR, U, G, threads, parts, slices ? &emit : nullptr);
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      best = std::min(best, s);
      bytes = emitted;
      for(auto& p : parts) bytes += p.size();
      if(slices && argc > 7 && i == 0) { FILE* f = fopen(argv[7], "w"); fwrite(sliced.data(), 1, sliced.size(), f); fclose(f); }
      if(i == 0 && getenv("MR_REPLAY_PARTS")) {          // how even the split between the threads is
        fprintf(stderr, "bytes per part:");
        for(auto& p : parts) fprintf(stderr, " %zu", p.size());
        fprintf(stderr, "\n");
      }
    }
    printf("{\"reads\": %u, \"rows\": %llu, \"text_bytes\": %llu, \"threads\": %u, \"best_s\": %.6f, \"thread_seconds\": %.6f}\n",
           d.view.nreads, (unsigned long long)d.view.ncoords, (unsigned long long)bytes, threads, best, best * threads);
    if(argc > 7 && !getenv("MR_REPLAY_SLICES")) { FILE* f = fopen(argv[7], "w"); for(auto& p : parts) fwrite(p.data(), 1, p.size(), f); fclose(f); }
  } catch(std::exception& e) { fprintf(stderr, "format_replay: %s\n", e.what()); return 1; }
  return 0;
}
