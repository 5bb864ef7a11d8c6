// Adaptive-seed ("adaptamer") merge of two sorted k-mer tables, fused with the seed-record
// transform of the re-import step.
//
// Replaces new_merge_thread (FastGA.c:610-1025) + adaptamer_merge (:2281-2493) and the per-pair
// arithmetic of reimport_thread (:2703-2721).  The reference co-walks T1 and a per-12-mer cache
// of T2 with an LCP-driven state machine; its result is the declarative rule:
//   for every forward-strand T1 entry e1 whose 12-base panel is non-empty in T2,
//     plen  = max over the panel of LCP(e1,e2)            (12..40)
//     R(e1) = the block of T2 entries with LCP(e1,e2) >= plen
//     if |R| < FREQ emit (plen, e1, e2) for every e2 in R, strand C iff e2 is a reverse entry.
// Here one thread owns one T1 entry: it binary-searches its 56-bit suffix in the T2 panel
// (located through the 2^24 prefix index), takes plen from the two neighbours of the insertion
// point, and walks at most FREQ entries either side.  T1 is streamed with coalesced 128-bit
// loads; the T2 panel of neighbouring threads is the same few cache lines, so T2 is read from
// HBM once.  Seeds of a block are compacted with a block scan and appended with one atomic.
#include "common.cuh"

typedef unsigned long long u64;

struct seed_layout                 // bit positions inside the 128-bit seed record
{ int anti_bits;                   // lcp [0,6) drem [6,12) anti [12,12+anti_bits) band ... jcont ... icont ... comp
  int band_bits;
  int jc_bits;
  int ic_bits;
  long long amxpos, bmxpos;        // longest contig of genome 1 / genome 2 (FastGA.c:5023-5041)
};

static __host__ __device__ __forceinline__ int seed_key_bits(const seed_layout &L)
{ return 12 + L.anti_bits + L.band_bits + L.jc_bits + L.ic_bits + 1; }

#define MG_THREADS 256
#define MG_WARPS   (MG_THREADS/32)
#define MG_TILE    64                     // T1 entries per warp (32 / 16 when T2 is much denser than T1)
#define MG_T2CAP   1280                   // staged T2 entries per block (20 KB)
#define MG_PCAP    1536                   // staged prefix-index entries per block (6 KB)

//  lcp (in bases, 0..28) of two 56-bit suffixes
static __device__ __forceinline__ int lcp56(u64 a, u64 b)
{ u64 x = a ^ b;
  return x ? ((__clzll(x) - 8) >> 1) : 28;
}

static __device__ __forceinline__ u64 suffix_of(const rec128 *__restrict__ T, unsigned i)
{ uint4 v = *reinterpret_cast<const uint4 *>(T + i);
  u64 lo = (u64) v.x | ((u64) v.y << 32);
  u64 hi = (u64) v.z | ((u64) v.w << 32);
  return ((hi & 0xffffffffffull) << 16) | (lo >> 48);
}

struct seed_pack                          // kernel-uniform packing constants
{ int p_band, s_jc, s_ic, s_cp;           // band position; shifts inside the upper word
  long long amxpos, bmxpos, maxdag;
};

struct blk_stage                          // the block's staging of T2
{ rec128 t2[MG_T2CAP];                    // the slice of T2 the block's tile can match (TMA destination),
                                          //   then rewritten in place as (56-bit suffix, payload) pairs
  unsigned ps[MG_PCAP];                   // the prefix-index range
  unsigned rng[4];
  unsigned wtot[MG_WARPS], wsum[MG_WARPS];
  unsigned long long gbase;
  unsigned long long bar;
};

struct warp_stage                         // one warp's private buffers
{ rec128 t1[MG_TILE];                     // forward-strand T1 entries of the warp's tile, compacted
  unsigned excl[MG_TILE+1], lowi[MG_TILE], plen[MG_TILE];
  u64    ipay[MG_TILE];
};

//  The two table views of the search: staged (shared memory, suffixes precomputed once per T2
//  entry) or direct (global memory, when a tile's slice does not fit the staging buffers).
struct view_staged
{ const blk_stage *S; unsigned t2off, psoff;
  __device__ __forceinline__ unsigned ps(unsigned p)  const { return S->ps[p - psoff]; }
  __device__ __forceinline__ u64 suf(unsigned i)      const { return reinterpret_cast<const u64 *>(S->t2)[2*(i - t2off)]; }
  __device__ __forceinline__ u64 pay(unsigned i)      const { return reinterpret_cast<const u64 *>(S->t2)[2*(i - t2off)+1]; }
};
struct view_direct
{ const rec128 *T2; const unsigned *pstart;
  __device__ __forceinline__ unsigned ps(unsigned p)  const { return pstart[p]; }
  __device__ __forceinline__ u64 suf(unsigned i)      const { return suffix_of(T2,i); }
  __device__ __forceinline__ u64 pay(unsigned i)      const { return T2[i].lo & 0xffffffffffffull; }
};

//  the adaptamer of one T1 entry: |R| (0 if no seed), first T2 index of R, plen
template<class View>
static __device__ __forceinline__ unsigned adaptamer(const View &V, const rec128 &r1, int freq,
                                                     unsigned &lowi, int &plen)
{ unsigned p  = KREC_PREFIX24(r1.hi);
  unsigned lo = V.ps(p), hi = V.ps(p+1);
  lowi = 0; plen = 0;
  if (lo >= hi) return 0;
  u64 s1 = KREC_SUFFIX56(r1);
  unsigned a = lo, b = hi;                                    // lower bound of s1 in T2[lo,hi)
  while (a < b)
    { unsigned m = (a + b) >> 1;
      if (V.suf(m) < s1) a = m+1; else b = m;
    }
  int ll = (a > lo) ? lcp56(s1,V.suf(a-1)) : -1;
  int lr = (a < hi) ? lcp56(s1,V.suf(a))   : -1;
  int m  = ll > lr ? ll : lr;
  plen = 12 + m;
  unsigned lft = a, rgt = a;
  int sh = 56 - 2*m;
  u64 key = s1 >> sh;
  while (lft > lo && rgt - lft < (unsigned) freq)
    { if ((V.suf(lft-1) >> sh) != key) break;
      lft -= 1;
    }
  while (rgt < hi && rgt - lft < (unsigned) freq)
    { if ((V.suf(rgt) >> sh) != key) break;
      rgt += 1;
    }
  if (rgt - lft >= (unsigned) freq) return 0;                 // |R| < FREQ (:799-823)
  lowi = lft;
  return rgt - lft;
}

//  One warp owns MG_TILE consecutive T1 entries; the block shares the staged slice of T2:
//   1. two coalesced 512-byte loads of the tile; forward-strand entries (the only ones that seed,
//      FastGA.c:921-928) are compacted into shared memory by ballot so the search lanes are dense;
//   2. the block's tiles are consecutive in k-mer order, so the T2 entries they can match are ONE
//      contiguous slice [pstart2[pA], pstart2[pB+1]): fetched with one TMA bulk copy together with
//      the prefix-index range, suffixes and payloads split once per staged entry;
//   3. each lane searches its entries in the slice (binary search + bounded walk);
//   4. the seeds of the tile are expanded load-balanced: output slot o is built by lane o mod 32
//      (prefix sums in shared memory), so every lane builds one seed per step and the 128-bit
//      stores of a step are contiguous; one atomic per tile reserves the output run.

template<class View>
static __device__ __forceinline__ unsigned merge_search(const View &V, warp_stage *S, int nd, int freq,
                                                        unsigned &sumlen, int lane)
{ unsigned run = 0, sl = 0;
  for (int r0 = 0; r0 < nd; r0 += 32)
    { int idx = r0 + lane;
      unsigned cnt = 0, lowi = 0; int plen = 0;
      u64 pay = 0;
      if (idx < nd)
        { rec128 r1 = ld_rec(&S->t1[idx]);
          cnt = adaptamer(V,r1,freq,lowi,plen);
          pay = r1.lo & 0xffffffffffffull;
        }
      unsigned inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
        { unsigned t = __shfl_up_sync(0xffffffffu,inc,o);
          if (lane >= o) inc += t;
        }
      if (idx < MG_TILE)
        { S->excl[idx] = run + inc - cnt; S->lowi[idx] = lowi; S->plen[idx] = (unsigned) plen; S->ipay[idx] = pay; }
      run += __shfl_sync(0xffffffffu,inc,31);
      sl  += cnt * (unsigned) plen;
    }
  const int nslot = (nd + 31) & ~31;                           // entries idx >= nd hold cnt 0
  if (lane == 0) S->excl[nslot] = run;
  sumlen = __reduce_add_sync(0xffffffffu,sl);
  __syncwarp();
  return run;
}

template<class View>
static __device__ __forceinline__ void merge_expand(const View &V, const warp_stage *S, int nd, unsigned total,
                                                    unsigned long long gbase, const seed_pack &K,
                                                    rec128 *__restrict__ seeds, unsigned long long capacity,
                                                    int lane)
{ const int nslot = (nd + 31) & ~31;
  for (unsigned o = lane; o < total; o += 32)
    { int j = 0;                                               // last slot with excl <= o
      for (int st = nslot >> 1; st > 0; st >>= 1)
        if (S->excl[j + st] <= o) j += st;
      unsigned k = o - S->excl[j];
      u64 p2 = V.pay(S->lowi[j] + k), pay = S->ipay[j];
      long long ipost = (long long) (unsigned) pay, jpost = (long long) (unsigned) p2;
      unsigned icont = (unsigned) (pay >> 32) & 0x7fff;
      unsigned cs = (unsigned) (p2 >> 32) & 0xffff;
      unsigned comp = cs >> 15, jcont = cs & 0x7fff;
      long long diag, anti;
      if (comp) { diag = K.maxdag - (ipost + jpost); anti = K.amxpos - (ipost - jpost); }
      else      { diag = K.bmxpos + (ipost - jpost); anti = ipost + jpost; }
      u64 X = (u64) S->plen[j] | ((u64) (diag & 63) << 6) | ((u64) anti << 12);
      u64 Y = (u64) (diag >> 6) | ((u64) jcont << K.s_jc) | ((u64) icont << K.s_ic)
                                | ((u64) comp << K.s_cp);
      rec128 sd;
      sd.lo = X | (Y << K.p_band);
      sd.hi = Y >> (64 - K.p_band);
      if (gbase + o < capacity) st_rec(seeds + gbase + o,sd);
    }
}

//  Per block of the merge: the prefix range of its T1 tile and the T2 slice it can match, found
//  ahead of the merge with full parallelism (inside the merge this is a chain of two dependent HBM
//  round trips made by one lane while 255 threads wait).
__global__ void merge_ranges_kernel(const rec128 *__restrict__ T1, unsigned n1, const unsigned *__restrict__ pstart2,
                                    unsigned per_block, unsigned nblocks, uint4 *__restrict__ rng)
{ unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  unsigned long long b0 = (unsigned long long) b * per_block, b1 = b0 + per_block - 1;
  if (b1 >= n1) b1 = n1 - 1;
  unsigned pA = KREC_PREFIX24(T1[b0].hi), pB = KREC_PREFIX24(T1[b1].hi);
  unsigned lo2 = pstart2[pA], hi2 = pstart2[pB+1];
  rng[b] = make_uint4(pA,pB - pA + 2,lo2,hi2 - lo2);
}

template<int TILE>
__global__ void __launch_bounds__(MG_THREADS)
adaptamer_merge_kernel(const rec128 *__restrict__ T1, unsigned n1,
                       const rec128 *__restrict__ T2, const unsigned *__restrict__ pstart2,
                       const uint4 *__restrict__ rng, int freq, seed_pack K,
                       rec128 *__restrict__ seeds, unsigned long long capacity,
                       unsigned long long *__restrict__ counters /* [0]=nseeds [1]=sum plen */)
{ extern __shared__ __align__(16) unsigned char mg_smem[];
  blk_stage  *B = reinterpret_cast<blk_stage *>(mg_smem);
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  warp_stage *S = reinterpret_cast<warp_stage *>(mg_smem + sizeof(blk_stage)) + wp;
  const unsigned lt = lanemask_lt();
  const unsigned long long b0 = (unsigned long long) blockIdx.x * MG_WARPS * TILE;
  const unsigned long long base = b0 + (unsigned long long) wp * TILE;

  //  every warp's tile loads are in flight before the block waits for its T2 slice
  unsigned i0 = (unsigned) base + lane, i1 = i0 + 32;
  rec128 e0, e1;
  e0.lo = e0.hi = e1.lo = e1.hi = 0;
  bool f0 = false, f1 = false;
  if (base < n1)
    { if (lane < TILE && i0 < n1) { e0 = ld_rec(T1 + i0); f0 = ((e0.lo >> 47) & 1) == 0; }
      if (TILE > 32 && i1 < n1)   { e1 = ld_rec(T1 + i1); f1 = ((e1.lo >> 47) & 1) == 0; }
    }

  //  The block's 512 T1 entries are consecutive in k-mer order, so the T2 entries they can match
  //  are one contiguous slice [pstart2[pA], pstart2[pB+1]).  When it fits, that slice and the
  //  prefix-index range are staged in shared memory (one TMA bulk copy + coalesced loads) and
  //  every search, walk and payload read hits shared memory instead of a dependent L2/HBM trip.
  if (threadIdx.x == 0)
    { const uint4 r = rng[blockIdx.x];                      // pA, #prefixes+1, first T2 entry, #T2 entries
      B->rng[0] = r.x; B->rng[1] = r.y; B->rng[2] = r.z; B->rng[3] = r.w;
      mbar_init(&B->bar,1);
      if (r.y <= MG_PCAP && r.w <= MG_T2CAP && r.w > 0)
        tma_load_1d(B->t2,T2 + r.z,r.w * 16u,&B->bar);
    }
  //  forward-strand compaction of the warp's tile while the slice is in flight
  unsigned m0 = __ballot_sync(0xffffffffu,f0), m1 = __ballot_sync(0xffffffffu,f1);
  int n0 = __popc(m0), nd = n0 + __popc(m1);
  if (f0) st_rec(&S->t1[__popc(m0 & lt)],e0);
  if (f1) st_rec(&S->t1[n0 + __popc(m1 & lt)],e1);
  __syncthreads();
  const unsigned pA = B->rng[0], nps = B->rng[1], lo2 = B->rng[2], nsl = B->rng[3];
  const bool staged = (nps <= MG_PCAP && nsl <= MG_T2CAP);
  if (staged)
    { for (unsigned i = threadIdx.x; i < nps; i += MG_THREADS) B->ps[i] = pstart2[pA + i];
      if (nsl > 0)
        { mbar_wait(&B->bar,0);
          for (unsigned j = threadIdx.x; j < nsl; j += MG_THREADS)
            { rec128 r = ld_rec(&B->t2[j]), c;
              c.lo = KREC_SUFFIX56(r);                       // first word: suffix, second: payload
              c.hi = r.lo & 0xffffffffffffull;
              st_rec(&B->t2[j],c);
            }
        }
    }
  __syncthreads();
  //  search, then ONE atomic per block reserves the output run of its eight tiles (a per-tile
  //  atomic on the single counter serialises in L2: 1.2 M same-address atomics cost > 1.5 ms)
  unsigned total = 0, sl = 0;
  view_staged Vs; Vs.S = B; Vs.t2off = lo2; Vs.psoff = pA;
  view_direct Vd; Vd.T2 = T2; Vd.pstart = pstart2;
  if (nd > 0)
    total = staged ? merge_search(Vs,S,nd,freq,sl,lane) : merge_search(Vd,S,nd,freq,sl,lane);
  if (lane == 0) { B->wtot[wp] = total; B->wsum[wp] = sl; }
  __syncthreads();
  if (threadIdx.x == 0)
    { unsigned long long t = 0, q = 0;
      for (int k = 0; k < MG_WARPS; k++) { t += B->wtot[k]; q += B->wsum[k]; }
      unsigned long long g = 0;
      if (t) { g = atomicAdd(&counters[0],t); atomicAdd(&counters[1],q); }
      B->gbase = g;
    }
  __syncthreads();
  if (total == 0) return;
  unsigned long long gbase = B->gbase;
  for (int k = 0; k < wp; k++) gbase += B->wtot[k];
  if (staged) merge_expand(Vs,S,nd,total,gbase,K,seeds,capacity,lane);
  else        merge_expand(Vd,S,nd,total,gbase,K,seeds,capacity,lane);
}

/***********************************************************************************************
 *  SELF mode (FastGA A): new_self_merge_thread (FastGA.c:1616-1909).  One table; EVERY entry,
 *  either strand, is an i.  plen = its longest prefix shared with another entry of its 12-base
 *  panel = max of the LCPs with its two neighbours; block = the run of entries sharing those plen
 *  bases, the entry included; if the block has < FREQ members, one seed (i, p) for every OTHER
 *  member p, strand C iff the signs differ.  One thread per entry: no search, its own index is
 *  the insertion point.  Output run of a block reserved with one atomic.
 **********************************************************************************************/

static __device__ __forceinline__ int lcp_full(const rec128 &a, const rec128 &b)      // bases, 0..40
{ u64 x = a.hi ^ b.hi;
  if (x) return __clzll(x) >> 1;
  unsigned y = (unsigned) ((a.lo ^ b.lo) >> 48);
  if (y) return 32 + ((__clz(y) - 16) >> 1);
  return 40;
}

__global__ void __launch_bounds__(256)
self_merge_kernel(const rec128 *__restrict__ T, unsigned n, const unsigned *__restrict__ pstart,
                  int freq, seed_pack K, rec128 *__restrict__ seeds, unsigned long long capacity,
                  unsigned long long *__restrict__ counters)
{ __shared__ unsigned s_w[8], s_l[8];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  unsigned lo = i, hi = i, cnt = 0;
  int plen = 0;
  rec128 e; e.lo = e.hi = 0;
  if (i < n)
    { e = ld_rec(T + i);
      unsigned p = KREC_PREFIX24(e.hi);
      unsigned cbeg = pstart[p], cend = pstart[p+1];
      int lp = (i > cbeg)   ? lcp_full(ld_rec(T + i - 1),e) : 11;
      int ls = (i+1 < cend) ? lcp_full(e,ld_rec(T + i + 1)) : 11;
      plen = lp > ls ? lp : ls;
      if (plen >= 12)
        { lo = i; hi = i+1;
          while (lo > cbeg && i - lo < (unsigned) freq && lcp_full(ld_rec(T + lo - 1),ld_rec(T + lo)) >= plen) lo -= 1;
          while (hi < cend && hi - lo < (unsigned) freq && lcp_full(ld_rec(T + hi - 1),ld_rec(T + hi)) >= plen) hi += 1;
          if (hi - lo < (unsigned) freq) cnt = hi - lo - 1;
        }
    }
  unsigned inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    { unsigned t = __shfl_up_sync(0xffffffffu,inc,o);
      if (lane >= o) inc += t;
    }
  unsigned slen = __reduce_add_sync(0xffffffffu,cnt * (unsigned) plen);
  if (lane == 31) { s_w[wp] = inc; s_l[wp] = slen; }
  __syncthreads();
  if (threadIdx.x == 0)
    { unsigned long long t = 0, q = 0;
      for (int k = 0; k < 8; k++) { t += s_w[k]; q += s_l[k]; }
      s_base = t ? atomicAdd(&counters[0],t) : 0ull;
      if (q) atomicAdd(&counters[1],q);
    }
  __syncthreads();
  if (cnt == 0) return;
  unsigned long long o = s_base + (inc - cnt);
  for (int k = 0; k < wp; k++) o += s_w[k];
  const long long ipost = (long long) (unsigned) e.lo;
  const unsigned icont = (unsigned) (e.lo >> 32) & 0x7fff, isign = (unsigned) (e.lo >> 47) & 1;
  for (unsigned q = lo; q < hi; q++)
    { if (q == i) continue;
      rec128 r2 = ld_rec(T + q);
      long long jpost = (long long) (unsigned) r2.lo;
      unsigned cs = (unsigned) (r2.lo >> 32) & 0xffff;
      unsigned comp = (cs >> 15) ^ isign, jcont = cs & 0x7fff;
      long long diag, anti;
      if (comp) { diag = K.maxdag - (ipost + jpost); anti = K.amxpos - (ipost - jpost); }
      else      { diag = K.bmxpos + (ipost - jpost); anti = ipost + jpost; }
      u64 X = (u64) (unsigned) plen | ((u64) (diag & 63) << 6) | ((u64) anti << 12);
      u64 Y = (u64) (diag >> 6) | ((u64) jcont << K.s_jc) | ((u64) icont << K.s_ic)
                                | ((u64) comp << K.s_cp);
      rec128 sd;
      sd.lo = X | (Y << K.p_band);
      sd.hi = Y >> (64 - K.p_band);
      if (o < capacity) st_rec(seeds + o,sd);
      o += 1;
    }
}

extern "C" int fgb_self_merge_device(const void *d_T, long long n, const unsigned *d_pstart, int freq,
                                     int anti_bits, int band_bits, int jc_bits, int ic_bits,
                                     long long amxpos, void *d_seeds, long long capacity,
                                     unsigned long long *d_counters, unsigned long long *h_nseeds,
                                     unsigned long long *h_sumlen, void *stream)
{ cudaStream_t st = (cudaStream_t) stream;
  seed_layout L;
  L.anti_bits = anti_bits; L.band_bits = band_bits; L.jc_bits = jc_bits; L.ic_bits = ic_bits;
  L.amxpos = amxpos; L.bmxpos = amxpos;
  if (seed_key_bits(L) > 128 || freq < 1 || freq > 255) return FGB_ERR_LIMIT;
  if (n >= 0xffffffffll) return FGB_ERR_LIMIT;
  CUDA_TRY(cudaMemsetAsync(d_counters,0,16,st));
  seed_pack K;
  K.p_band = 12 + anti_bits;
  K.s_jc = band_bits; K.s_ic = band_bits + jc_bits; K.s_cp = band_bits + jc_bits + ic_bits;
  K.amxpos = amxpos; K.bmxpos = amxpos; K.maxdag = 2*amxpos;
  if (K.s_cp + 1 > 64 || K.p_band >= 64 || K.p_band < 13) return FGB_ERR_LIMIT;
  if (n > 0)
    { self_merge_kernel<<<(unsigned) ((n + 255) / 256),256,0,st>>>((const rec128 *) d_T,(unsigned) n,d_pstart,freq,K,
                                                                  (rec128 *) d_seeds,(unsigned long long) capacity,d_counters);
      fgb_count_launch(1);
    }
  CUDA_TRY(cudaGetLastError());
  unsigned long long h[2];
  CUDA_TRY(cudaMemcpyAsync(h,d_counters,16,cudaMemcpyDeviceToHost,st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *h_nseeds = h[0];
  if (h_sumlen) *h_sumlen = h[1];
  return (h[0] > (unsigned long long) capacity) ? FGB_ERR_OVERFLOW : FGB_OK;
}

//  T1/T2: sorted device tables; pstart2: [2^24+1] lower-bound index of T2.  Appends seed
//  records to d_seeds (capacity records).  d_counters: 2 x u64 on the device, zeroed here.
//  On return *h_nseeds is the number of seeds FOUND; if it exceeds capacity the buffer
//  content is incomplete and the caller must retry with a larger buffer (FGB_ERR_OVERFLOW).

extern "C" int fgb_merge_device(const void *d_T1, long long n1, const void *d_T2, long long n2,
                                const unsigned *d_pstart2, int freq,
                                int anti_bits, int band_bits, int jc_bits, int ic_bits,
                                long long amxpos, long long bmxpos,
                                void *d_seeds, long long capacity, unsigned long long *d_counters,
                                unsigned long long *h_nseeds, unsigned long long *h_sumlen,
                                void *stream)
{ cudaStream_t st = (cudaStream_t) stream;
  seed_layout L;
  L.anti_bits = anti_bits; L.band_bits = band_bits; L.jc_bits = jc_bits; L.ic_bits = ic_bits;
  L.amxpos = amxpos; L.bmxpos = bmxpos;
  if (seed_key_bits(L) > 128 || freq < 1 || freq > 255) return FGB_ERR_LIMIT;
  if (n1 >= 0xffffffffll) return FGB_ERR_LIMIT;
  CUDA_TRY(cudaMemsetAsync(d_counters,0,16,st));
  seed_pack K;
  K.p_band = 12 + anti_bits;
  K.s_jc = band_bits; K.s_ic = band_bits + jc_bits; K.s_cp = band_bits + jc_bits + ic_bits;
  K.amxpos = amxpos; K.bmxpos = bmxpos; K.maxdag = amxpos + bmxpos;
  if (K.s_cp + 1 > 64 || K.p_band >= 64 || K.p_band < 13) return FGB_ERR_LIMIT;
  if (n1 > 0)
    { //  T2 denser than T1 (a shard of genome 1 against all of genome 2): smaller tiles keep the
      //  block's T2 slice inside the staging buffer
      int tile = MG_TILE;
      if (n2 > 0 && n1 > 0)
        { double ratio = (double) n2 / (double) n1;
          if (ratio > 3.2) tile = 16; else if (ratio > 1.6) tile = 32;
        }
      unsigned nb = (unsigned) ((n1 + MG_WARPS*tile - 1) / (MG_WARPS*tile));
      cudaEvent_t ea, eb;
      cudaEventCreate(&ea); cudaEventCreate(&eb);
      static bool attr_set = false;
      const int smem = (int) (sizeof(blk_stage) + MG_WARPS*sizeof(warp_stage));
      if (!attr_set)
        { CUDA_TRY(cudaFuncSetAttribute(adaptamer_merge_kernel<64>,cudaFuncAttributeMaxDynamicSharedMemorySize,smem));
          CUDA_TRY(cudaFuncSetAttribute(adaptamer_merge_kernel<32>,cudaFuncAttributeMaxDynamicSharedMemorySize,smem));
          CUDA_TRY(cudaFuncSetAttribute(adaptamer_merge_kernel<16>,cudaFuncAttributeMaxDynamicSharedMemorySize,smem));
          attr_set = true;
        }
      uint4 *d_rng = NULL;
      CUDA_TRY(fgb_dmalloc((void **) &d_rng,sizeof(uint4)*(size_t) nb,st));
      cudaEventRecord(ea,st);
      merge_ranges_kernel<<<(nb + 255)/256,256,0,st>>>((const rec128 *) d_T1,(unsigned) n1,d_pstart2,
                                                      (unsigned) (MG_WARPS*tile),nb,d_rng);
#define MG_LAUNCH(T) adaptamer_merge_kernel<T><<<nb,MG_THREADS,smem,st>>>((const rec128 *) d_T1,(unsigned) n1, \
                       (const rec128 *) d_T2,d_pstart2,d_rng,freq,K,(rec128 *) d_seeds,(unsigned long long) capacity,d_counters)
      if (tile == 64) MG_LAUNCH(64); else if (tile == 32) MG_LAUNCH(32); else MG_LAUNCH(16);
#undef MG_LAUNCH
      cudaEventRecord(eb,st);
      cudaEventSynchronize(eb);
      float ms = 0; cudaEventElapsedTime(&ms,ea,eb);
      fgb_timing_add(3,ms);
      fgb_count_launch(2);
      cudaEventDestroy(ea); cudaEventDestroy(eb);
      fgb_dfree(d_rng,st);
    }
  CUDA_TRY(cudaGetLastError());
  unsigned long long h[2];
  CUDA_TRY(cudaMemcpyAsync(h,d_counters,16,cudaMemcpyDeviceToHost,st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *h_nseeds = h[0];
  if (h_sumlen) *h_sumlen = h[1];
  return (h[0] > (unsigned long long) capacity) ? FGB_ERR_OVERFLOW : FGB_OK;
}
