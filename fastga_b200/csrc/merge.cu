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
#define MG_TILE    64                     // T1 entries per warp
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

//  One warp owns MG_TILE consecutive T1 entries:
//   1. two coalesced 512-byte loads; forward-strand entries (the only ones that seed,
//      FastGA.c:921-928) are compacted into shared memory by ballot so the search lanes are dense;
//   2. each lane searches its entry in the T2 panel (binary search + bounded walk);
//   3. the seeds of the 32 lanes are expanded load-balanced: output slot o is built by lane
//      o mod 32 (prefix sums in shared memory), so every lane builds one seed per step and the
//      128-bit stores of a step are contiguous.

__global__ void __launch_bounds__(MG_THREADS)
adaptamer_merge_kernel(const rec128 *__restrict__ T1, unsigned n1,
                       const rec128 *__restrict__ T2, const unsigned *__restrict__ pstart2,
                       int freq, seed_pack K,
                       rec128 *__restrict__ seeds, unsigned long long capacity,
                       unsigned long long *__restrict__ counters /* [0]=nseeds [1]=sum plen */)
{ __shared__ __align__(16) rec128 sbuf[MG_WARPS][MG_TILE];
  __shared__ __align__(16) rec128 s_T2[MG_T2CAP];        // the block's slice of T2 ...
  __shared__ unsigned s_ps[MG_PCAP];                      // ... and of its prefix index
  __shared__ unsigned s_rng[4];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ unsigned s_excl[MG_WARPS][33];
  __shared__ unsigned s_lowi[MG_WARPS][32];
  __shared__ unsigned s_plen[MG_WARPS][32];
  __shared__ unsigned long long s_pay[MG_WARPS][32];
  __shared__ unsigned long long s_sum[MG_WARPS];

  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  unsigned long long base = ((unsigned long long) blockIdx.x * MG_WARPS + wp) * MG_TILE;
  unsigned long long sumlen = 0;

  //  The block's 512 T1 entries are consecutive in k-mer order, so the T2 entries they can match
  //  are one contiguous slice [pstart2[pA], pstart2[pB+1]).  When it fits, that slice and the
  //  prefix-index range are staged in shared memory with coalesced loads and every search,
  //  walk and payload read below hits shared memory instead of a dependent L2/HBM round trip.
  { unsigned long long b0 = (unsigned long long) blockIdx.x * MG_WARPS * MG_TILE;
    if (threadIdx.x == 0)
      { unsigned long long b1 = b0 + MG_WARPS*MG_TILE - 1;
        if (b1 >= n1) b1 = n1 - 1;
        unsigned pA = KREC_PREFIX24(T1[b0].hi), pB = KREC_PREFIX24(T1[b1].hi);
        unsigned lo2 = pstart2[pA], hi2 = pstart2[pB+1];
        s_rng[0] = pA; s_rng[1] = pB - pA + 2; s_rng[2] = lo2; s_rng[3] = hi2 - lo2;
        mbar_init(&s_bar,1);
        //  the T2 slice is one contiguous run of 128-bit records: a single TMA bulk copy
        if (pB - pA + 2 <= MG_PCAP && hi2 - lo2 <= MG_T2CAP && hi2 > lo2)
          tma_load_1d(s_T2,T2 + lo2,(hi2 - lo2) * 16u,&s_bar);
      }
    __syncthreads();
  }
  const unsigned pA = s_rng[0], lo2 = s_rng[2];
  const bool staged = (s_rng[1] <= MG_PCAP && s_rng[3] <= MG_T2CAP);
  if (staged)
    { for (unsigned i = threadIdx.x; i < s_rng[1]; i += MG_THREADS) s_ps[i] = pstart2[pA + i];
      if (s_rng[3] > 0) mbar_wait(&s_bar,0);
    }
  __syncthreads();
  const rec128   *T2v = staged ? s_T2 : T2;     const unsigned t2off = staged ? lo2 : 0;
  const unsigned *psv = staged ? s_ps : pstart2; const unsigned psoff = staged ? pA  : 0;

  if (base < n1)
    { unsigned i0 = (unsigned) base + lane, i1 = i0 + 32;
      rec128 e0, e1;
      bool f0 = false, f1 = false;
      if (i0 < n1) { e0 = ld_rec(T1 + i0); f0 = ((e0.lo >> 47) & 1) == 0; }
      if (i1 < n1) { e1 = ld_rec(T1 + i1); f1 = ((e1.lo >> 47) & 1) == 0; }
      unsigned m0 = __ballot_sync(0xffffffffu,f0), m1 = __ballot_sync(0xffffffffu,f1);
      int n0 = __popc(m0), nd = n0 + __popc(m1);
      if (f0) st_rec(&sbuf[wp][__popc(m0 & lt)],e0);
      if (f1) st_rec(&sbuf[wp][n0 + __popc(m1 & lt)],e1);
      __syncwarp();

      for (int r0 = 0; r0 < nd; r0 += 32)
        { int idx = r0 + lane;
          unsigned cnt = 0, lowi = 0;
          int plen = 0;
          rec128 r1; r1.lo = r1.hi = 0;
          if (idx < nd)
            { r1 = ld_rec(&sbuf[wp][idx]);
              unsigned p  = KREC_PREFIX24(r1.hi);
              unsigned lo = psv[p - psoff], hi = psv[p + 1 - psoff];
              if (lo < hi)
                { u64 s1 = KREC_SUFFIX56(r1);
                  unsigned a = lo, b = hi;                    // lower bound of s1 in T2[lo,hi)
                  while (a < b)
                    { unsigned m = (a + b) >> 1;
                      if (suffix_of(T2v,m - t2off) < s1) a = m+1; else b = m;
                    }
                  int ll = (a > lo) ? lcp56(s1,suffix_of(T2v,a-1 - t2off)) : -1;
                  int lr = (a < hi) ? lcp56(s1,suffix_of(T2v,a - t2off))   : -1;
                  int m  = ll > lr ? ll : lr;
                  plen = 12 + m;
                  unsigned lft = a, rgt = a;
                  int sh = 56 - 2*m;
                  u64 key = s1 >> sh;
                  while (lft > lo && rgt - lft < (unsigned) freq)
                    { if ((suffix_of(T2v,lft-1 - t2off) >> sh) != key) break;
                      lft -= 1;
                    }
                  while (rgt < hi && rgt - lft < (unsigned) freq)
                    { if ((suffix_of(T2v,rgt - t2off) >> sh) != key) break;
                      rgt += 1;
                    }
                  if (rgt - lft < (unsigned) freq)             // |R| < FREQ (:799-823)
                    { cnt = rgt - lft; lowi = lft; }
                }
            }

          unsigned inc = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
            { unsigned t = __shfl_up_sync(0xffffffffu,inc,o);
              if (lane >= o) inc += t;
            }
          unsigned total = __shfl_sync(0xffffffffu,inc,31);
          if (total == 0) continue;
          sumlen += (unsigned long long) __reduce_add_sync(0xffffffffu,cnt * (unsigned) plen);
          unsigned long long gbase = 0;
          if (lane == 0) gbase = atomicAdd(&counters[0],(unsigned long long) total);
          gbase = __shfl_sync(0xffffffffu,gbase,0);
          s_excl[wp][lane] = inc - cnt;
          if (lane == 31) s_excl[wp][32] = total;
          s_lowi[wp][lane] = lowi;
          s_plen[wp][lane] = (unsigned) plen;
          s_pay[wp][lane]  = r1.lo & 0xffffffffffffull;
          __syncwarp();

          for (unsigned o = lane; o < total; o += 32)
            { int j = 0;                                       // last lane with s_excl <= o
#pragma unroll
              for (int st = 16; st > 0; st >>= 1)
                if (s_excl[wp][j + st] <= o) j += st;
              unsigned k = o - s_excl[wp][j];
              rec128 r2 = ld_rec(T2v + (s_lowi[wp][j] + k - t2off));
              unsigned long long pay = s_pay[wp][j];
              long long ipost = (long long) (unsigned) pay, jpost = (long long) (unsigned) r2.lo;
              unsigned icont = (unsigned) (pay >> 32) & 0x7fff;
              unsigned cs = (unsigned) (r2.lo >> 32) & 0xffff;
              unsigned comp = cs >> 15, jcont = cs & 0x7fff;
              long long diag, anti;
              if (comp) { diag = K.maxdag - (ipost + jpost); anti = K.amxpos - (ipost - jpost); }
              else      { diag = K.bmxpos + (ipost - jpost); anti = ipost + jpost; }
              u64 X = (u64) s_plen[wp][j] | ((u64) (diag & 63) << 6) | ((u64) anti << 12);
              u64 Y = (u64) (diag >> 6) | ((u64) jcont << K.s_jc) | ((u64) icont << K.s_ic)
                                        | ((u64) comp << K.s_cp);
              rec128 sd;
              sd.lo = X | (Y << K.p_band);
              sd.hi = Y >> (64 - K.p_band);
              if (gbase + o < capacity) st_rec(seeds + gbase + o,sd);
            }
          __syncwarp();
        }
    }

  if (lane == 0) s_sum[wp] = sumlen;
  __syncthreads();
  if (threadIdx.x == 0)
    { unsigned long long t = 0;
      for (int k = 0; k < MG_WARPS; k++) t += s_sum[k];
      if (t) atomicAdd(&counters[1],t);
    }
}

//  T1/T2: sorted device tables; pstart2: [2^24+1] lower-bound index of T2.  Appends seed
//  records to d_seeds (capacity records).  d_counters: 2 x u64 on the device, zeroed here.
//  On return *h_nseeds is the number of seeds FOUND; if it exceeds capacity the buffer
//  content is incomplete and the caller must retry with a larger buffer (FGB_ERR_OVERFLOW).

extern "C" int fgb_merge_device(const void *d_T1, long long n1, const void *d_T2,
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
    { unsigned nb = (unsigned) ((n1 + MG_WARPS*MG_TILE - 1) / (MG_WARPS*MG_TILE));
      cudaEvent_t ea, eb;
      cudaEventCreate(&ea); cudaEventCreate(&eb);
      cudaEventRecord(ea,st);
      adaptamer_merge_kernel<<<nb,MG_THREADS,0,st>>>((const rec128 *) d_T1,(unsigned) n1,
                                                     (const rec128 *) d_T2,d_pstart2,freq,K,
                                                     (rec128 *) d_seeds,(unsigned long long) capacity,
                                                     d_counters);
      cudaEventRecord(eb,st);
      cudaEventSynchronize(eb);
      float ms = 0; cudaEventElapsedTime(&ms,ea,eb);
      fgb_timing_add(3,ms);
      fgb_count_launch(1);
      cudaEventDestroy(ea); cudaEventDestroy(eb);
    }
  CUDA_TRY(cudaGetLastError());
  unsigned long long h[2];
  CUDA_TRY(cudaMemcpyAsync(h,d_counters,16,cudaMemcpyDeviceToHost,st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *h_nseeds = h[0];
  if (h_sumlen) *h_sumlen = h[1];
  return (h[0] > (unsigned long long) capacity) ? FGB_ERR_OVERFLOW : FGB_OK;
}
