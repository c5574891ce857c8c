// Index serialisation: the reference rebuilds its suffix array in every process, and the
// pipeline runs create_mega_reads several times over the same super-reads
// (mega_reads_assemble_cluster2.sh:358-448 array jobs; passes over the same -r in
// mega_reads_assemble_nomatch.sh:217-261).  A built index is a handful of flat device arrays, so it
// is written to / read from one file: a fixed header, then the arrays back to back.  The header
// carries a checksum of the INPUTS (text, super-read starts, unitig tables, psa-min, mer) so that a
// caller can tell whether a file on disk belongs to the inputs it has just parsed.
#include "index.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

int finish_single_part(mr_index* idx, const std::vector<uint32_t>& sr_len);   // index.cu
int build_slots(mr_index* idx);                                                 // index.cu

namespace {

// version 2: reserved[2] = hash of the part's section bytes (verified by mr_index_load), reserved[3] = layout
// tag derived from the parameters the arrays depend on (mi, tail width, block shift)
constexpr char     kMagic[8] = { 'M', 'R', 'B', '2', 'I', 'D', 'X', '2' };
constexpr uint32_t kSections = 9;     // text sa tails counts sr_start blk unitig_ids unitig_off unitig_len

struct file_header {
  char     magic[8];
  uint64_t n;
  uint32_t nsa, nseq, k, m, mi, tail_bits, tail_bytes, nshort, n_unitigs, has_unitigs;
  uint64_t short_key[kMaxShort];
  uint64_t bytes[kSections];
  uint64_t inputs_checksum;
  uint64_t reserved[4];
};

inline uint64_t mix(uint64_t h, uint64_t v) {      // splitmix64 finaliser over a running value
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ULL;
  h ^= h >> 27; h *= 0x94d049bb133111ebULL;
  h ^= h >> 31;
  return h;
}

// running hash of the section bytes as they pass through the staging buffer (8 bytes per step, a
// multiply and a rotate: several GB/s, far above the file read)
inline uint64_t body_hash(uint64_t h, const void* p, size_t len) {
  const unsigned char* b = (const unsigned char*)p;
  size_t i = 0;
  for(; i + 8 <= len; i += 8) {
    uint64_t w;
    memcpy(&w, b + i, 8);
    h = (h ^ w) * 0x9e3779b97f4a7c15ULL;
    h = (h << 29) | (h >> 35);
  }
  uint64_t w = 0;
  if(i < len) { memcpy(&w, b + i, len - i); h = (h ^ w) * 0x9e3779b97f4a7c15ULL; h = (h << 29) | (h >> 35); }
  return mix(h, len);
}
inline uint64_t layout_tag(uint32_t mi, uint32_t tail_bits, uint32_t tail_bytes) {
  return mix(mix(mix(mix(0x6c61796f757432ULL, mi), tail_bits), tail_bytes), (uint64_t)kBlkShift);
}

struct section { dev_buf* buf; uint64_t bytes; };

void sections_of(mr_index* idx, section* s) {
  const uint64_t nwords = (idx->n + 31) / 32;
  const uint32_t nprefix = 1u << (2 * idx->mi);
  const uint32_t nblk = (uint32_t)(idx->n >> kBlkShift) + 1;
  s[0] = { &idx->text,     (nwords + 2) * sizeof(uint64_t) };
  s[1] = { &idx->sa,       (uint64_t)idx->nsa * sizeof(uint32_t) };
  s[2] = { &idx->tails,    ((uint64_t)idx->nsa + 64) * idx->view.tail_bytes };
  s[3] = { &idx->counts,   ((uint64_t)nprefix + 8) * sizeof(uint32_t) };
  s[4] = { &idx->sr_start, ((uint64_t)idx->nseq + 2) * sizeof(uint32_t) };
  s[5] = { &idx->blk,      ((uint64_t)nblk + 1) * sizeof(uint32_t) };
  s[6] = { &idx->unitig_ids, 0 }; s[7] = { &idx->unitig_off, 0 }; s[8] = { &idx->unitig_len, 0 };
  if(idx->has_unitigs) {
    s[6].bytes = idx->unitig_total * sizeof(uint32_t);
    s[7].bytes = ((uint64_t)(idx->nseq_all ? idx->nseq_all : idx->nseq) + 1) * sizeof(uint64_t);   // spans all parts
    s[8].bytes = (uint64_t)idx->n_unitigs * sizeof(int32_t);
  }
}

struct file_closer { void operator()(FILE* f) const { if(f) fclose(f); } };
constexpr size_t kChunk = 64u << 20;

} // namespace

extern "C" {

uint64_t mr_inputs_checksum(const uint64_t* text2bit, uint64_t n, const uint64_t* sr_start, uint32_t nseq,
                            const uint32_t* unitig_ids, const uint64_t* unitig_off, const int32_t* unitig_len,
                            uint32_t n_unitigs, uint32_t psa_min, uint32_t k) {
  uint64_t h = mix(mix(mix(mix(0x6d72623230306964ULL, n), nseq), psa_min), k);
  if(text2bit) {
    const uint64_t nwords = n / 32;
    for(uint64_t i = 0; i < nwords; ++i) h = mix(h, text2bit[i]);
    if(n & 31) h = mix(h, text2bit[nwords] & ((1ULL << (2 * (n & 31))) - 1));   // bits past the last base are not part of the input
  }
  if(sr_start) for(uint32_t i = 0; i <= nseq; ++i) h = mix(h, sr_start[i]);
  if(unitig_ids && unitig_off && unitig_len && n_unitigs) {
    h = mix(h, n_unitigs);
    for(uint32_t i = 0; i <= nseq; ++i) h = mix(h, unitig_off[i]);
    for(uint64_t i = 0; i < unitig_off[nseq]; ++i) h = mix(h, unitig_ids[i]);
    for(uint32_t i = 0; i < n_unitigs; ++i) h = mix(h, (uint64_t)(uint32_t)unitig_len[i]);
  }
  return h;
}

uint64_t mr_index_checksum(const mr_index* idx) { return idx ? idx->inputs_checksum : 0; }

} // extern "C"

namespace {

// one part: header + its sections.  Whole-index fields of the header: reserved[0] = number of parts
// (written in the first header only), reserved[1] = sr_base of the part.
int save_part(mr_context* ctx, mr_index* idx, FILE* f, pinned_buf& stage, uint32_t nparts) {
  file_header h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kMagic, 8);
  h.n = idx->n; h.nsa = idx->nsa; h.nseq = idx->nseq; h.k = idx->k; h.m = idx->m; h.mi = idx->mi;
  h.tail_bits = idx->view.tail_bits; h.tail_bytes = idx->view.tail_bytes; h.nshort = idx->view.nshort;
  h.n_unitigs = idx->n_unitigs; h.has_unitigs = idx->has_unitigs;
  memcpy(h.short_key, idx->view.short_key, sizeof h.short_key);
  h.inputs_checksum = idx->inputs_checksum;
  h.reserved[0] = nparts; h.reserved[1] = idx->view.sr_base;
  section s[kSections];
  sections_of(idx, s);
  for(uint32_t i = 0; i < kSections; ++i) h.bytes[i] = s[i].bytes;
  h.reserved[3] = layout_tag(idx->mi, idx->view.tail_bits, idx->view.tail_bytes);
  const long header_at = ftell(f);
  if(header_at < 0 || fwrite(&h, sizeof h, 1, f) != 1) return ctx->fail(MR_EINVAL, "mr_index_save: write error");
  uint64_t bh = 0x6d72626f6479ULL;
  for(uint32_t i = 0; i < kSections; ++i) {
    for(uint64_t off = 0; off < s[i].bytes; off += kChunk) {
      const size_t len = (size_t)std::min<uint64_t>(kChunk, s[i].bytes - off);
      MR_CUDA(ctx, cudaMemcpyAsync(stage.p, (const char*)s[i].buf->p + off, len, cudaMemcpyDeviceToHost, ctx->stream));
      MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      bh = body_hash(bh, stage.p, len);
      if(fwrite(stage.p, 1, len, f) != len) return ctx->fail(MR_EINVAL, "mr_index_save: write error");
    }
  }
  // the hash of the body is only known now: rewrite the header in place
  h.reserved[2] = bh;
  const long end_at = ftell(f);
  if(end_at < 0 || fseek(f, header_at, SEEK_SET) != 0 || fwrite(&h, sizeof h, 1, f) != 1 || fseek(f, end_at, SEEK_SET) != 0)
    return ctx->fail(MR_EINVAL, "mr_index_save: write error");
  return MR_OK;
}

int load_part(mr_context* ctx, FILE* f, pinned_buf* stage, int& which, mr_index* idx, file_header& h, std::vector<uint32_t>& sr_len) {
  if(fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0)
    return ctx->fail(MR_EINVAL, "mr_index_load: not an index file of this library version");
  if(!(h.m >= 1 && h.m < h.k && h.k <= 31 && h.mi >= 1 && h.mi < h.k && h.n >= h.k && h.n < 0xfffffff0ULL && h.nseq >= 1 &&
       h.nsa == (uint32_t)(h.n - h.m + 1) && h.tail_bits == 2 * (h.k - h.mi) && h.nshort <= (uint32_t)kMaxShort &&
       (h.tail_bytes == 1 || h.tail_bytes == 2 || h.tail_bytes == 4)))
    return ctx->fail(MR_EINVAL, "mr_index_load: inconsistent header");
  if(h.reserved[3] != layout_tag(h.mi, h.tail_bits, h.tail_bytes))
    return ctx->fail(MR_EINVAL, "mr_index_load: the file was written with another array layout");
  // the unitig sections are the only ones whose size is not implied by the fields above: bound them
  // (a path entry per super-read base at most; unitig_off spans the super-reads of all parts, each of
  // which has at least one base of a text of fewer than 2^34)
  if(h.has_unitigs && (h.bytes[6] % sizeof(uint32_t) != 0 || h.bytes[6] > (1ULL << 36) || h.bytes[7] % sizeof(uint64_t) != 0 ||
                       h.bytes[7] < ((uint64_t)h.nseq + 1) * sizeof(uint64_t) || h.bytes[7] > ((1ULL << 34) + 1) * sizeof(uint64_t) ||
                       h.n_unitigs == 0))
    return ctx->fail(MR_EINVAL, "mr_index_load: inconsistent header (unitig sections)");
  idx->ctx = ctx; idx->n = h.n; idx->nsa = h.nsa; idx->nseq = h.nseq; idx->k = h.k; idx->m = h.m; idx->mi = h.mi;
  idx->n_unitigs = h.n_unitigs; idx->has_unitigs = h.has_unitigs != 0;
  idx->inputs_checksum = h.inputs_checksum;
  idx->view.tail_bytes = h.tail_bytes;
  if(idx->has_unitigs) idx->unitig_total = h.bytes[6] / sizeof(uint32_t);
  section s[kSections];
  sections_of(idx, s);
  if(idx->has_unitigs) s[7].bytes = h.bytes[7];      // unitig_off spans the super-reads of ALL parts
  for(uint32_t i = 0; i < kSections; ++i)
    if(s[i].bytes != h.bytes[i]) return ctx->fail(MR_EINVAL, "mr_index_load: section sizes do not match the header");
  MR_TRY(idx->alloc_lut(ctx, s[3].bytes, s[2].bytes));
  uint64_t bh = 0x6d72626f6479ULL;
  uint32_t max_uid = 0;
  for(uint32_t i = 0; i < kSections; ++i) {
    if(s[i].bytes == 0) continue;
    MR_TRY(s[i].buf->ensure(ctx, s[i].bytes));
    for(uint64_t off = 0; off < s[i].bytes; off += kChunk) {
      const size_t len = (size_t)std::min<uint64_t>(kChunk, s[i].bytes - off);
      // two staging buffers: the file read of one chunk overlaps the upload of the previous one
      MR_CUDA(ctx, cudaEventSynchronize(ctx->ev[which]));
      if(fread(stage[which].p, 1, len, f) != len) return ctx->fail(MR_EINVAL, "mr_index_load: file is truncated");
      bh = body_hash(bh, stage[which].p, len);
      if(i == 6) {                               // largest unitig id of the paths: see mr_index::unitig_ids_ok
        const uint32_t* ids = (const uint32_t*)stage[which].p;
        for(size_t q = 0; q < len / sizeof(uint32_t); ++q) max_uid = std::max(max_uid, ids[q] >> 1);
      }
      MR_CUDA(ctx, cudaMemcpyAsync((char*)s[i].buf->p + off, stage[which].p, len, cudaMemcpyHostToDevice, ctx->stream));
      MR_CUDA(ctx, cudaEventRecord(ctx->ev[which], ctx->stream));
      which ^= 1;
    }
  }
  MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if(bh != h.reserved[2]) return ctx->fail(MR_EINVAL, "mr_index_load: the arrays do not match the hash in the header (corrupt or stale file)");
  idx->unitig_ids_ok = !idx->has_unitigs || idx->unitig_total == 0 || max_uid < idx->n_unitigs;
  index_view& v = idx->view;
  v.counts = idx->counts.as<uint32_t>(); v.tails = idx->tails.p; v.sa = idx->sa.as<uint32_t>();
  v.sr_start = idx->sr_start.as<uint32_t>(); v.blk = idx->blk.as<uint32_t>();
  v.n = h.n; v.nsa = h.nsa; v.nseq = h.nseq; v.k = h.k; v.m = h.m; v.mi = h.mi; v.tail_bits = h.tail_bits; v.tail_bytes = h.tail_bytes;
  v.nshort = h.nshort;
  v.sr_base = (uint32_t)h.reserved[1]; v.nseq_all = h.nseq;
  v.own = 0;                                     // set below from the loaded starts
  memcpy(v.short_key, h.short_key, sizeof h.short_key);
  idx->n_all = h.n; idx->nseq_all = h.nseq;
  {                                              // super-read lengths from the starts just loaded
    std::vector<uint32_t> st(h.nseq + 1);
    MR_CUDA(ctx, cudaMemcpy(st.data(), idx->sr_start.p, ((size_t)h.nseq + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for(uint32_t i = 0; i < h.nseq; ++i) sr_len.push_back(st[i + 1] - st[i]);
    v.own = st[h.nseq];
  }
  MR_TRY(build_slots(idx));                      // tables derived from the loaded arrays: not part of the file
  return MR_OK;
}

} // namespace

extern "C" {

int mr_index_save(mr_index* idx, const char* path) {
  if(!idx || !path) return MR_EINVAL;
  mr_context* ctx = idx->ctx;
  MR_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<FILE, file_closer> f(fopen(path, "wb"));
  if(!f) return ctx->fail(MR_EINVAL, std::string("mr_index_save: cannot open ") + path);
  pinned_buf stage;
  MR_TRY(stage.ensure(ctx, kChunk));
  MR_TRY(save_part(ctx, idx, f.get(), stage, idx->nparts()));
  for(mr_index* part : idx->more) MR_TRY(save_part(ctx, part, f.get(), stage, 0));
  if(fflush(f.get()) != 0) return ctx->fail(MR_EINVAL, "mr_index_save: write error");
  return MR_OK;
}

int mr_index_peek_checksum(const char* path, uint64_t* checksum) {
  if(!path || !checksum) return MR_EINVAL;
  std::unique_ptr<FILE, file_closer> f(fopen(path, "rb"));
  file_header h;
  if(!f || fread(&h, sizeof h, 1, f.get()) != 1 || memcmp(h.magic, kMagic, 8) != 0) return MR_EINVAL;
  *checksum = h.inputs_checksum;
  return MR_OK;
}

int mr_index_load(mr_context* ctx, const char* path, mr_index** out) {
  if(!ctx) return MR_EINVAL;
  if(!path || !out) return ctx->fail(MR_EINVAL, "mr_index_load: null argument");
  *out = nullptr;
  MR_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<FILE, file_closer> f(fopen(path, "rb"));
  if(!f) return ctx->fail(MR_EINVAL, std::string("mr_index_load: cannot open ") + path);
  ctx->timers.clear();
  phase_timer timer(ctx);
  timer.begin("index load");
  pinned_buf stage[2];
  MR_TRY(stage[0].ensure(ctx, kChunk)); MR_TRY(stage[1].ensure(ctx, kChunk));
  int which = 0;
  std::unique_ptr<mr_index> idx(new mr_index);
  std::vector<uint32_t> sr_len;
  file_header h;
  MR_TRY(load_part(ctx, f.get(), stage, which, idx.get(), h, sr_len));
  const uint32_t nparts = h.reserved[0] ? (uint32_t)h.reserved[0] : 1u;
  if(nparts > (uint32_t)kMaxParts) return ctx->fail(MR_EINVAL, "mr_index_load: inconsistent header");
  uint64_t n_all = h.n;
  for(uint32_t p = 1; p < nparts; ++p) {
    std::unique_ptr<mr_index> part(new mr_index);
    file_header hp;
    MR_TRY(load_part(ctx, f.get(), stage, which, part.get(), hp, sr_len));
    if(hp.k != h.k || hp.m != h.m || part->view.sr_base != sr_len.size() - hp.nseq)
      return ctx->fail(MR_EINVAL, "mr_index_load: parts do not fit together");
    n_all += hp.n;
    idx->more.push_back(part.release());
  }
  timer.end();
  timer.collect();
  const uint32_t nseq_all = (uint32_t)sr_len.size();
  idx->nseq_all = nseq_all;
  idx->view.nseq_all = nseq_all;
  for(mr_index* part : idx->more) { part->nseq_all = nseq_all; part->view.nseq_all = nseq_all; }
  if(nparts > 1) {
    // a part's text carries the first k-1 bases of the next part (index.cuh): they are not its own
    n_all = 0;
    for(uint32_t i = 0; i < nseq_all; ++i) n_all += sr_len[i];
    std::vector<index_view> views(nparts - 1);
    for(uint32_t p = 1; p < nparts; ++p) views[p - 1] = idx->more[p - 1]->view;
    MR_TRY(idx->more_views.ensure(ctx, views.size() * sizeof(index_view)));
    MR_CUDA(ctx, cudaMemcpy(idx->more_views.p, views.data(), views.size() * sizeof(index_view), cudaMemcpyHostToDevice));
  }
  idx->n_all = n_all;
  MR_TRY(finish_single_part(idx.get(), sr_len));
  *out = idx.release();
  return MR_OK;
}

} // extern "C"
