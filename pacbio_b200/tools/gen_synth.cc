// Synthetic-input generator for the create_mega_reads hot path:
// random genome (optionally repeat-rich) -> k-unitigs -> named super-reads ->
// simulated PacBio reads.  Deterministic for a given seed, independent of the
// number of threads.  Output files are the ones the reference CLI takes:
//   <prefix>.unitigs.fa       one-line FASTA, record index == unitig id  (-u)
//   <prefix>.unitigs_len.txt  "id len" lines                             (-l)
//   <prefix>.superreads.fa    headers are unitig paths "12F_7R_..."      (-r)
//   <prefix>.reads.fa         simulated long reads                       (-p)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

struct rng_t {  // xoshiro256** seeded through splitmix64
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  explicit rng_t(uint64_t seed) { for(auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
  double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

static inline char comp(char c) {
  switch(c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}
static std::string revcomp(const char* s, size_t n) {
  std::string r(n, 'N');
  for(size_t i = 0; i < n; ++i) r[i] = comp(s[n - 1 - i]);
  return r;
}

struct options {
  uint64_t genome = 1000000, seed = 42;
  double   coverage = 20, error = 0.15, sr_cov = 2.0, repeat_frac = 0.0, single_frac = 0.1;
  uint32_t read_len = 10000, unitig_k = 41, mean_unitig = 500, threads = 8, line = 80;
  uint32_t shards = 1;     // > 1: also <prefix>.reads.shard<i>.fa, i = 1 .. shards-1, each as many NEW reads as reads.fa
  uint32_t first_shard = 0; // > 0: write only the read shards first_shard .. shards-1 (the other files exist already)
  std::string prefix = "synth";
};

int main(int argc, char** argv) {
  options o;
  for(int i = 1; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    const char* v = argv[i + 1];
    if(a == "--genome") o.genome = strtoull(v, 0, 0);
    else if(a == "--seed") o.seed = strtoull(v, 0, 0);
    else if(a == "--coverage") o.coverage = atof(v);
    else if(a == "--error") o.error = atof(v);
    else if(a == "--sr-cov") o.sr_cov = atof(v);
    else if(a == "--repeat-frac") o.repeat_frac = atof(v);
    else if(a == "--single-frac") o.single_frac = atof(v);
    else if(a == "--read-len") o.read_len = atoi(v);
    else if(a == "--unitig-k") o.unitig_k = atoi(v);
    else if(a == "--mean-unitig") o.mean_unitig = atoi(v);
    else if(a == "--threads") o.threads = atoi(v);
    else if(a == "--shards") o.shards = std::max(1, atoi(v));
    else if(a == "--first-shard") o.first_shard = std::max(0, atoi(v));
    else if(a == "--prefix") o.prefix = v;
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
  }
  const uint32_t K = o.unitig_k;
  const uint64_t G = o.genome;
  rng_t rng(o.seed);

  // ---- genome -------------------------------------------------------------
  std::string genome(G, 'A');
  for(uint64_t i = 0; i < G; ) {
    uint64_t r = rng.next();
    for(int j = 0; j < 32 && i < G; ++j, ++i, r >>= 2) genome[i] = "ACGT"[r & 3];
  }
  // repeat families: exact copies pasted over the random background.  Copies are
  // kept disjoint so that every copy's interior is exactly one shared unitig.
  struct rep_copy { uint64_t pos; uint32_t len; uint32_t family; bool exact; };
  std::vector<rep_copy> copies;
  if(o.repeat_frac > 0 && G > 100000) {
    std::vector<uint8_t> used((G >> 6) + 2, 0);  // 64-base granularity occupancy
    uint64_t covered = 0;
    uint32_t family = 0;
    while(covered < (uint64_t)(o.repeat_frac * G)) {
      const uint32_t len = 300 + rng.below(5700);
      uint64_t ncopies = 10;
      { const double u = rng.unit(); ncopies = (uint64_t)(10 * std::min(1000.0, 1.0 / (u + 1e-3))); } // 10..10^4, heavy tail
      ncopies = std::min<uint64_t>(ncopies, std::max<uint64_t>(2, (uint64_t)(o.repeat_frac * G / 20 / len)));
      const double diverge = rng.unit() < 0.5 ? 0.0 : rng.unit() * 0.05;
      std::string unit(len, 'A');
      for(auto& c : unit) c = "ACGT"[rng.below(4)];
      for(uint64_t c = 0; c < ncopies; ++c) {
        uint64_t pos = 0; bool ok = false;
        for(int tries = 0; tries < 20 && !ok; ++tries) {
          pos = K + rng.below(G - len - 2 * K);
          ok = true;
          for(uint64_t b = (pos - K) >> 6; b <= (pos + len + K) >> 6; ++b) if(used[b]) { ok = false; break; }
        }
        if(!ok) continue;
        for(uint64_t b = (pos - K) >> 6; b <= (pos + len + K) >> 6; ++b) used[b] = 1;
        std::string cp = unit;
        if(diverge > 0) for(auto& ch : cp) if(rng.unit() < diverge) ch = "ACGT"[rng.below(4)];
        memcpy(&genome[pos], cp.data(), len);
        copies.push_back({ pos, len, family, diverge == 0.0 });
        covered += len;
      }
      ++family;
    }
    std::sort(copies.begin(), copies.end(), [](const rep_copy& a, const rep_copy& b) { return a.pos < b.pos; });
  }

  size_t n_unitigs_out = 0, nseg_out = 0;
  uint64_t sr_bases = 0, nsr = 0;
  if(o.first_shard == 0) {
  // ---- unitigs: segment i = genome[cut[i], cut[i+1] + K - 1) --------------
  std::vector<uint64_t> cut;
  std::vector<int64_t>  seg_family;   // family id if the segment is an exact repeat interior, else -1
  {
    size_t ci = 0;
    uint64_t pos = 0;
    cut.push_back(0); seg_family.push_back(-1);
    while(true) {
      uint64_t next = pos + std::max<uint64_t>(K / 2 + 1, (uint64_t)(o.mean_unitig * (0.25 + 1.5 * rng.unit())));
      while(ci < copies.size() && copies[ci].pos + copies[ci].len <= pos + K) ++ci;
      if(ci < copies.size() && next > copies[ci].pos && copies[ci].len > 2 * K) {
        const rep_copy& c = copies[ci];
        if(c.pos > pos) { cut.push_back(c.pos); seg_family.push_back(c.exact ? (int64_t)c.family : -1); }
        else seg_family.back() = c.exact ? (int64_t)c.family : -1;
        next = c.pos + c.len - (K - 1);
        ++ci;
      }
      if(next + K >= G) break;
      cut.push_back(next); seg_family.push_back(-1);
      pos = next;
    }
  }
  const size_t nseg = cut.size();
  auto seg_end = [&](size_t i) { return i + 1 < nseg ? cut[i + 1] + K - 1 : G; };
  // unitig ids: exact repeat interiors of one family share an id
  std::vector<uint32_t> seg_uid(nseg);
  std::vector<uint8_t>  seg_flip(nseg);   // stored unitig is the reverse complement of the genome segment
  std::vector<size_t>   uid_seg;          // representative segment for each unitig id
  std::vector<uint8_t>  uid_flip;
  {
    std::vector<int64_t> family_uid;
    for(size_t i = 0; i < nseg; ++i) {
      const int64_t f = seg_family[i];
      if(f >= 0) {
        if((size_t)f >= family_uid.size()) family_uid.resize(f + 1, -1);
        if(family_uid[f] >= 0) { seg_uid[i] = family_uid[f]; seg_flip[i] = uid_flip[seg_uid[i]]; continue; }
        family_uid[f] = uid_seg.size();
      }
      seg_uid[i] = uid_seg.size();
      uid_seg.push_back(i);
      uid_flip.push_back(rng.below(2));
      seg_flip[i] = uid_flip.back();
    }
  }
  {
    FILE* fa = fopen((o.prefix + ".unitigs.fa").c_str(), "w");
    FILE* fl = fopen((o.prefix + ".unitigs_len.txt").c_str(), "w");
    if(!fa || !fl) { perror("open"); return 1; }
    for(size_t u = 0; u < uid_seg.size(); ++u) {
      const size_t i = uid_seg[u];
      const uint64_t s = cut[i], e = seg_end(i);
      fprintf(fa, ">%zu\n", u);
      if(uid_flip[u]) { const std::string r = revcomp(&genome[s], e - s); fwrite(r.data(), 1, r.size(), fa); }
      else fwrite(&genome[s], 1, e - s, fa);
      fputc('\n', fa);
      fprintf(fl, "%zu %llu\n", u, (unsigned long long)(e - s));
    }
    fclose(fa); fclose(fl);
  }

  n_unitigs_out = uid_seg.size(); nseg_out = nseg;
  // ---- super-reads ---------------------------------------------------------
  {
    FILE* f = fopen((o.prefix + ".superreads.fa").c_str(), "w");
    if(!f) { perror("open"); return 1; }
    std::unordered_set<std::string> seen;
    const uint64_t target = (uint64_t)(o.sr_cov * G);
    uint64_t fails = 0;
    while(sr_bases < target && fails < 1000000) {
      const size_t len = rng.unit() < o.single_frac ? 1 : 2 + rng.below(11);
      if(len > nseg) { ++fails; continue; }
      const size_t first = rng.below(nseg - len + 1);
      const bool   rev   = rng.below(2);
      std::string name;
      for(size_t t = 0; t < len; ++t) {
        const size_t i = rev ? first + len - 1 - t : first + t;
        if(t) name += '_';
        name += std::to_string(seg_uid[i]);
        name += (seg_flip[i] ^ rev) ? 'R' : 'F';
      }
      // one orientation per unitig path: the same run in both orientations would give two
      // super-reads aligning at identical read coordinates, an exact tie that the reference
      // itself orders nondeterministically (unordered_map + unstable sort)
      std::string rname;
      for(size_t t = 0; t < len; ++t) {
        const size_t i = rev ? first + t : first + len - 1 - t;
        if(t) rname += '_';
        rname += std::to_string(seg_uid[i]);
        rname += (seg_flip[i] ^ rev) ? 'F' : 'R';
      }
      if(!seen.insert(std::min(name, rname)).second) { ++fails; continue; }
      const uint64_t s = cut[first], e = seg_end(first + len - 1);
      std::string seq = rev ? revcomp(&genome[s], e - s) : genome.substr(s, e - s);
      fprintf(f, ">%s\n", name.c_str());
      for(size_t p = 0; p < seq.size(); p += o.line) {
        fwrite(seq.data() + p, 1, std::min<size_t>(o.line, seq.size() - p), f);
        fputc('\n', f);
      }
      sr_bases += seq.size(); ++nsr;
    }
    fclose(f);
  }
  }   // first_shard == 0

  // ---- reads ---------------------------------------------------------------
  const uint64_t nreads = std::max<uint64_t>(1, (uint64_t)(o.coverage * G / o.read_len));
  uint64_t read_bases = 0;
  // shard s holds reads [s * nreads, (s + 1) * nreads): disjoint draws from the same genome (a read's
  // random stream depends on its index only), so shard 0 is the file a run without --shards writes
  for(uint32_t shard = o.first_shard; shard < o.shards; ++shard) {
    FILE* f = fopen((o.prefix + (shard ? ".reads.shard" + std::to_string(shard) + ".fa" : std::string(".reads.fa"))).c_str(), "w");
    if(!f) { perror("open"); return 1; }
    const uint64_t chunk = 4096;
    std::vector<std::string> bufs(o.threads);
    std::vector<uint64_t> nb(o.threads);
    const uint64_t first_read = (uint64_t)shard * nreads, end_read = first_read + nreads;
    for(uint64_t base = first_read; base < end_read; base += chunk * o.threads) {
      std::vector<std::thread> th;
      for(uint32_t t = 0; t < o.threads; ++t) {
        th.emplace_back([&, t]() {
          std::string& out = bufs[t];
          out.clear(); nb[t] = 0;
          const uint64_t lo = base + t * chunk, hi = std::min(end_read, lo + chunk);
          for(uint64_t r = lo; r < hi; ++r) {
            rng_t g(o.seed * 0x100000001b3ULL + r + 1);
            uint64_t len = (uint64_t)(o.read_len * (0.8 + 0.4 * g.unit()));
            len = std::min(len, G);
            const uint64_t start = g.below(G - len + 1);
            const bool rev = g.below(2);
            std::string src = rev ? revcomp(&genome[start], len) : genome.substr(start, len);
            char hdr[128];
            snprintf(hdr, sizeof(hdr), ">read%llu/%llu_%llu_%c\n", (unsigned long long)r,
                     (unsigned long long)start, (unsigned long long)(start + len), rev ? '-' : '+');
            out += hdr;
            const size_t before = out.size();
            for(uint64_t i = 0; i < len; ++i) {
              if(g.unit() < o.error) {
                const double u = g.unit();
                if(u < 0.5) { out += "ACGT"[g.below(4)]; out += src[i]; }        // insertion
                else if(u < 0.8) { /* deletion */ }
                else { char c; do { c = "ACGT"[g.below(4)]; } while(c == src[i]); out += c; }
              } else out += src[i];
            }
            if(g.below(1000) == 0 && out.size() - before > 100) {               // an N-run in 1 read per 1000
              const size_t rl = out.size() - before, nlen = 1 + g.below(20), at = g.below(rl - nlen);
              for(size_t i = 0; i < nlen; ++i) out[before + at + i] = 'N';
            }
            nb[t] += out.size() - before;
            out += '\n';
          }
        });
      }
      for(auto& x : th) x.join();
      for(uint32_t t = 0; t < o.threads; ++t) { fwrite(bufs[t].data(), 1, bufs[t].size(), f); if(shard == o.first_shard) read_bases += nb[t]; }
    }
    fclose(f);
  }
  printf("{\"genome\": %llu, \"unitigs\": %zu, \"segments\": %zu, \"superreads\": %llu, \"superread_bases\": %llu, "
         "\"reads\": %llu, \"read_bases\": %llu, \"unitig_k\": %u}\n",
         (unsigned long long)G, n_unitigs_out, nseg_out, (unsigned long long)nsr, (unsigned long long)sr_bases,
         (unsigned long long)nreads, (unsigned long long)read_bases, K);
  return 0;
}
