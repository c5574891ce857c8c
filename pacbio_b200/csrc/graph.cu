// Overlap graph between the super-reads aligned to one read: nodes ordered by implied start,
// O(n^2) edge test (position overlap vs unitig-path dovetail overlap), longest-path DP and
// union-find components.  One warp per read (graph_kernel) or, for a read with many rows, one CTA
// (graph_big_kernel); the outer node loop is sequential as in the reference, the inner loop over
// candidate successors runs 32 (256) wide with a ballot for the reference's `break`.  Replaces overlap_graph::thread::reset + overlap_graph::traverse
// (overlap_graph.hpp:24-34,177-196, overlap_graph.cc:7-59, union_find.cc:6-24,
// super_read_name.cc:49-72).
#include "align.cuh"

namespace {

struct path_ref {
  const uint32_t* ids; uint32_t n; bool bwd;
  __device__ uint32_t at(uint32_t t) const { return bwd ? (ids[n - 1 - t] ^ 1u) : ids[t]; }
};

__device__ __forceinline__ path_ref row_path(const graph_args& A, uint64_t row) {
  path_ref p;
  if(!A.unitig_off) { p.ids = nullptr; p.n = 0; p.bwd = false; return p; }
  const uint32_t sr = A.c.sr[row];
  const uint64_t u0 = A.unitig_off[sr];
  p.ids = A.unitig_ids + u0;
  p.n   = (uint32_t)(A.unitig_off[sr + 1] - u0);
  p.bwd = A.c.use_bwd[row] != 0;
  return p;
}

// largest t such that the last t unitigs of l equal the first t of r (super_read_name.cc:49-72)
__device__ int dovetail(const path_ref& l, const path_ref& r) {
  if(l.n < 2 || r.n < 2) return 0;
  int32_t first = (int32_t)l.n - (int32_t)r.n + 1;
  if(first < 1) first = 1;
  const uint32_t r0 = r.at(0);
  for(uint32_t i = (uint32_t)first; i < l.n; ++i) {
    if(l.at(i) != r0) continue;
    uint32_t j = i + 1;
    while(j < l.n && l.at(j) == r.at(j - i)) ++j;
    if(j == l.n) return (int)(l.n - i);
  }
  return 0;
}

__device__ bool same_path(const path_ref& a, const path_ref& b) {
  if(a.n != b.n) return false;
  for(uint32_t t = 0; t < a.n; ++t) if(a.at(t) != b.at(t)) return false;
  return true;
}

__device__ int uf_find(int32_t* parent, int s) {
  int r = s;
  while(parent[r] != r) r = parent[r];
  while(parent[s] != r) { const int nx = parent[s]; parent[s] = r; s = nx; }
  return r;
}

__global__ void __launch_bounds__(128) graph_kernel(graph_args A) {
  __shared__ int32_t edge_j[4][32];
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if(r >= A.nreads) return;
  const uint64_t b = A.read_coords[r];
  const int n = (int)(A.read_coords[r + 1] - b);
  if(n == 0 || n > A.warp_max_rows) return;           // many rows: graph_big_kernel
  const double rl = (double)A.read_len[r];
  const double K = (double)A.unitigs_k;
  int32_t* parent = A.component + b;
  int32_t* rank   = A.uf_rank + b;
  int32_t* order  = A.order + b;
  double*  imp_s  = A.imp_s + b;
  double*  imp_e  = A.imp_e + b;

  // node_info::reset (overlap_graph.hpp:24-34)
  for(int i = lane; i < n; i += 32) {
    const uint64_t row = b + i;
    const double st = A.c.stretch[row], of = A.c.offset[row];
    imp_s[i] = st + of;
    const double t = st * (double)A.c.ql[row];
    imp_e[i] = t + of;
    A.start_node[row] = 1; A.end_node[row] = 1;
    parent[i] = i; rank[i] = 0;
    A.lstart[row] = -1; A.lprev[row] = -1;
    A.lpath[row] = A.bases ? (int32_t)A.c.sr_cover[row] : A.c.nb_mers[row];
    A.lunitigs[row] = (int32_t)row_path(A, row).n;
  }
  __syncwarp();
  // node order by (imp_s, imp_e); exact ties keep row order (the reference's std::sort is unstable there)
  for(int i = lane; i < n; i += 32) {
    const double s = imp_s[i], e = imp_e[i];
    int rk = 0;
    for(int j = 0; j < n; ++j) {
      const double sj = imp_s[j], ej = imp_e[j];
      rk += (sj < s || (sj == s && ej < e)) || (sj == s && ej == e && j < i);
    }
    order[rk] = i;
  }
  __syncwarp();

  for(int a = 0; a < n; ++a) {
    const int ii = order[a];
    const uint64_t row_i = b + ii;
    const double ie_i = imp_e[ii];
    if(ie_i >= rl) continue;                         // hanging off the 3' end of the read
    const path_ref pi = row_path(A, row_i);
    const double err_i = A.c.avg_err[row_i];
    const int lpath_i = A.lpath[row_i], lstart_i = A.lstart[row_i], lunitigs_i = A.lunitigs[row_i];
    const double start_s_i = imp_s[lstart_i == -1 ? ii : lstart_i];
    bool any_edge = false;
    for(int b0 = a + 1; b0 < n; b0 += 32) {
      const int bb = b0 + (int)lane;
      const bool in = bb < n;
      const int jj = in ? order[bb] : 0;
      const uint64_t row_j = b + jj;
      const double is_j = imp_s[jj], ie_j = imp_e[jj];
      const bool skip = !in || is_j <= 1 || ie_i > ie_j + 31;
      const double position_len = ie_i - is_j;
      const double error1 = err_i + A.c.avg_err[row_j];
      const double error  = A.errors * error1;
      const double ppl = position_len * A.overlap_play;
      const bool brk = !skip && (ppl + error < K);
      const unsigned ball = __ballot_sync(MR_FULL_MASK, brk);
      const unsigned limit = ball ? (unsigned)(__ffs(ball) - 1) : 32u;
      bool edge = false;
      int nb_u = 0, common = 0;
      path_ref pj; pj.ids = nullptr; pj.n = 0; pj.bwd = false;
      if(!skip && lane < limit) {
        pj = row_path(A, row_j);
        nb_u = dovetail(pi, pj);
        if(nb_u && !same_path(pi, pj)) {
          int u_overlap_len = 0;
          const uint32_t ilen = A.c.info_len[row_j];
          const int32_t* info = (A.bases ? A.binfo : A.kinfo) + A.c.info_off[row_j];
          for(int u = 0; u < nb_u; ++u) {
            u_overlap_len += A.unitig_len[pj.at(u) >> 1];
            if((uint32_t)(2 * u) < ilen) common += info[2 * u];
            if(u > 0 && (uint32_t)(2 * u - 1) < ilen) common -= info[2 * u - 1];
          }
          u_overlap_len -= (nb_u - 1) * ((int)A.unitigs_k - 1);
          const double t1 = A.overlap_play * position_len;
          const double t2 = A.overlap_play * ((double)u_overlap_len + error);
          edge = !((double)u_overlap_len > t1 + error || position_len > t2);
        }
      }
      const unsigned eb = __ballot_sync(MR_FULL_MASK, edge);
      if(eb) {
        any_edge = true;
        if(edge) {
          A.start_node[row_j] = 0;
          const int nlpath = lpath_i + (A.bases ? (int)A.c.sr_cover[row_j] : A.c.nb_mers[row_j]) - common;
          const int lpath_j = A.lpath[row_j], lstart_j = A.lstart[row_j];
          const double start_s_j = imp_s[lstart_j == -1 ? jj : lstart_j];
          if(nlpath > lpath_j || (nlpath == lpath_j && (lstart_j == -1 || start_s_i > start_s_j))) {
            A.lpath[row_j]    = nlpath;
            A.lstart[row_j]   = lstart_i == -1 ? ii : lstart_i;
            A.lprev[row_j]    = ii;
            A.lunitigs[row_j] = lunitigs_i + (int)pj.n - nb_u;
          }
        }
        // components: unions in successor order, exactly as the sequential loop would do them
        edge_j[wib][lane] = jj;
        __syncwarp();
        if(lane == 0) {
          unsigned m = eb;
          while(m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int r1 = uf_find(parent, ii), r2 = uf_find(parent, edge_j[wib][src]);
            if(rank[r1] > rank[r2]) parent[r2] = r1;
            else if(rank[r1] < rank[r2]) parent[r1] = r2;
            else if(r1 != r2) { parent[r2] = r1; ++rank[r1]; }
          }
        }
        __syncwarp();
      }
      if(ball) break;
    }
    if(any_edge && lane == 0) A.end_node[row_i] = 0;
    __syncwarp();
  }

  // component root of every node (union_find::set::root)
  for(int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + (int)lane;
    int root = 0;
    if(i < n) { root = i; while(parent[root] != root) root = parent[root]; }
    __syncwarp();
    if(i < n) parent[i] = root;
    __syncwarp();
  }
}

// The same for a read with many rows (repeats: hundreds to thousands): one CTA per read, of 256 threads
// up to cta_max_rows rows and of 1024 above (the scan of a node's successors takes rows / threads rounds).  The outer
// node loop stays sequential; the candidate successors are tested kBigThreads at a time, the
// reference's `break` becomes the first flagged thread of the block, and the unions of a round are
// still done by one thread in successor order.
template<int kBigThreads>
__global__ void __launch_bounds__(kBigThreads) graph_big_kernel(graph_args A, int min_rows, int max_rows) {
  __shared__ int32_t  edge_j[kBigThreads];
  __shared__ unsigned s_brk[kBigThreads / 32], s_edge[kBigThreads / 32];
  const int tid = (int)threadIdx.x;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kBigThreads / 32;
  const uint32_t r = blockIdx.x;
  if(r >= A.nreads) return;
  const uint64_t b = A.read_coords[r];
  const int n = (int)(A.read_coords[r + 1] - b);
  if(n <= min_rows || n > max_rows) return;
  const double rl = (double)A.read_len[r];
  const double K = (double)A.unitigs_k;
  int32_t* parent = A.component + b;
  int32_t* rank   = A.uf_rank + b;
  int32_t* order  = A.order + b;
  double*  imp_s  = A.imp_s + b;
  double*  imp_e  = A.imp_e + b;
  // the same three values per node in node ORDER: the successor scan reads them without going through order[]
  double*  ord_s   = A.ord_s + b;
  double*  ord_e   = A.ord_e + b;
  double*  ord_err = A.ord_err + b;

  for(int i = tid; i < n; i += kBigThreads) {
    const uint64_t row = b + i;
    const double st = A.c.stretch[row], of = A.c.offset[row];
    imp_s[i] = st + of;
    const double t = st * (double)A.c.ql[row];
    imp_e[i] = t + of;
    A.start_node[row] = 1; A.end_node[row] = 1;
    parent[i] = i; rank[i] = 0;
    A.lstart[row] = -1; A.lprev[row] = -1;
    A.lpath[row] = A.bases ? (int32_t)A.c.sr_cover[row] : A.c.nb_mers[row];
    A.lunitigs[row] = (int32_t)row_path(A, row).n;
  }
  __syncthreads();
  for(int i = tid; i < n; i += kBigThreads) {
    const double s = imp_s[i], e = imp_e[i];
    int rk = 0;
    for(int j = 0; j < n; ++j) {
      const double sj = imp_s[j], ej = imp_e[j];
      rk += (sj < s || (sj == s && ej < e)) || (sj == s && ej == e && j < i);
    }
    order[rk] = i;
    ord_s[rk] = s; ord_e[rk] = e; ord_err[rk] = A.c.avg_err[b + i];
  }
  __syncthreads();

  for(int a = 0; a < n; ++a) {
    const int ii = order[a];
    const uint64_t row_i = b + ii;
    const double ie_i = imp_e[ii];
    if(ie_i >= rl) continue;                         // uniform over the block
    const path_ref pi = row_path(A, row_i);
    const double err_i = A.c.avg_err[row_i];
    const int lpath_i = A.lpath[row_i], lstart_i = A.lstart[row_i], lunitigs_i = A.lunitigs[row_i];
    const double start_s_i = imp_s[lstart_i == -1 ? ii : lstart_i];
    bool any_edge = false;
    for(int b0 = a + 1; b0 < n; b0 += kBigThreads) {
      const int bb = b0 + tid;
      const bool in = bb < n;
      const int jj = in ? order[bb] : 0;
      const uint64_t row_j = b + jj;
      const double is_j = in ? ord_s[bb] : 0.0, ie_j = in ? ord_e[bb] : 0.0;
      const bool skip = !in || is_j <= 1 || ie_i > ie_j + 31;
      const double position_len = ie_i - is_j;
      const double error1 = err_i + (in ? ord_err[bb] : 0.0);
      const double error  = A.errors * error1;
      const double ppl = position_len * A.overlap_play;
      const bool brk = !skip && (ppl + error < K);
      const unsigned ball = __ballot_sync(MR_FULL_MASK, brk);
      if(lane == 0) s_brk[warp] = ball;
      __syncthreads();
      int limit = kBigThreads;                       // first thread of the block that breaks
      for(int w = 0; w < NW; ++w) if(s_brk[w]) { limit = w * 32 + __ffs(s_brk[w]) - 1; break; }
      bool edge = false;
      int nb_u = 0, common = 0;
      path_ref pj; pj.ids = nullptr; pj.n = 0; pj.bwd = false;
      if(!skip && tid < limit) {
        pj = row_path(A, row_j);
        nb_u = dovetail(pi, pj);
        if(nb_u && !same_path(pi, pj)) {
          int u_overlap_len = 0;
          const uint32_t ilen = A.c.info_len[row_j];
          const int32_t* info = (A.bases ? A.binfo : A.kinfo) + A.c.info_off[row_j];
          for(int u = 0; u < nb_u; ++u) {
            u_overlap_len += A.unitig_len[pj.at(u) >> 1];
            if((uint32_t)(2 * u) < ilen) common += info[2 * u];
            if(u > 0 && (uint32_t)(2 * u - 1) < ilen) common -= info[2 * u - 1];
          }
          u_overlap_len -= (nb_u - 1) * ((int)A.unitigs_k - 1);
          const double t1 = A.overlap_play * position_len;
          const double t2 = A.overlap_play * ((double)u_overlap_len + error);
          edge = !((double)u_overlap_len > t1 + error || position_len > t2);
        }
      }
      const unsigned eb = __ballot_sync(MR_FULL_MASK, edge);
      if(lane == 0) s_edge[warp] = eb;
      edge_j[tid] = jj;
      __syncthreads();
      unsigned any = 0;
      for(int w = 0; w < NW; ++w) any |= s_edge[w];
      if(any) {
        any_edge = true;
        if(edge) {
          A.start_node[row_j] = 0;
          const int nlpath = lpath_i + (A.bases ? (int)A.c.sr_cover[row_j] : A.c.nb_mers[row_j]) - common;
          const int lpath_j = A.lpath[row_j], lstart_j = A.lstart[row_j];
          const double start_s_j = imp_s[lstart_j == -1 ? jj : lstart_j];
          if(nlpath > lpath_j || (nlpath == lpath_j && (lstart_j == -1 || start_s_i > start_s_j))) {
            A.lpath[row_j]    = nlpath;
            A.lstart[row_j]   = lstart_i == -1 ? ii : lstart_i;
            A.lprev[row_j]    = ii;
            A.lunitigs[row_j] = lunitigs_i + (int)pj.n - nb_u;
          }
        }
        if(tid == 0) {                               // unions in successor order, as the sequential loop does them
          for(int w = 0; w < NW; ++w) {
            unsigned m = s_edge[w];
            while(m) {
              const int src = __ffs(m) - 1;
              m &= m - 1;
              const int r1 = uf_find(parent, ii), r2 = uf_find(parent, edge_j[w * 32 + src]);
              if(rank[r1] > rank[r2]) parent[r2] = r1;
              else if(rank[r1] < rank[r2]) parent[r1] = r2;
              else if(r1 != r2) { parent[r2] = r1; ++rank[r1]; }
            }
          }
        }
      }
      __syncthreads();                               // shared arrays are rewritten by the next round
      if(limit < kBigThreads) break;
    }
    if(any_edge && tid == 0) A.end_node[row_i] = 0;
    __syncthreads();                                 // the rows updated in this step are read by the next one
  }

  for(int i0 = 0; i0 < n; i0 += kBigThreads) {
    const int i = i0 + tid;
    int root = 0;
    if(i < n) { root = i; while(parent[root] != root) root = parent[root]; }
    __syncthreads();
    if(i < n) parent[i] = root;
    __syncthreads();
  }
}

} // namespace

int launch_graph(mr_context* ctx, const graph_args& a) {
  if(a.nreads == 0) return MR_OK;
  graph_kernel<<<div_up((uint64_t)a.nreads * 32, 128), 128, 0, ctx->stream>>>(a);
  MR_LAUNCHED(ctx);
  // reads with many rows: a CTA each (the kernels return at once for every other read)
  graph_big_kernel<256><<<a.nreads, 256, 0, ctx->stream>>>(a, a.warp_max_rows, a.cta_max_rows);
  MR_LAUNCHED(ctx);
  graph_big_kernel<1024><<<a.nreads, 1024, 0, ctx->stream>>>(a, a.cta_max_rows, 0x7fffffff);
  MR_LAUNCHED(ctx);
  return MR_OK;
}
