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

static __device__ __forceinline__ void put_bits(rec128 &r, int pos, u64 v)
{ if (pos < 64)
    { r.lo |= v << pos;
      if (pos > 0) r.hi |= v >> (64-pos);     // spill (v is narrower than 64 bits)
    }
  else
    r.hi |= v << (pos-64);
}

static __device__ __forceinline__ rec128 make_seed(const seed_layout &L, int comp, unsigned icont,
                                                   unsigned jcont, long long ipost, long long jpost,
                                                   int plen)
{ long long diag, anti;
  if (comp)                                        // FastGA.c:2705-2712
    { diag = (L.amxpos + L.bmxpos) - (ipost + jpost);
      anti = L.amxpos - (ipost - jpost);
    }
  else
    { diag = L.bmxpos + (ipost - jpost);
      anti = ipost + jpost;
    }
  rec128 r; r.lo = 0; r.hi = 0;
  int pos = 0;
  put_bits(r,pos,(u64) plen);            pos += 6;
  put_bits(r,pos,(u64) (diag & 63));     pos += 6;
  put_bits(r,pos,(u64) anti);            pos += L.anti_bits;
  put_bits(r,pos,(u64) (diag >> 6));     pos += L.band_bits;
  put_bits(r,pos,(u64) jcont);           pos += L.jc_bits;
  put_bits(r,pos,(u64) icont);           pos += L.ic_bits;
  put_bits(r,pos,(u64) comp);
  return r;
}

#define MG_THREADS 256

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

__global__ void __launch_bounds__(MG_THREADS)
adaptamer_merge_kernel(const rec128 *__restrict__ T1, unsigned n1,
                       const rec128 *__restrict__ T2, const unsigned *__restrict__ pstart2,
                       int freq, seed_layout L,
                       rec128 *__restrict__ seeds, unsigned long long capacity,
                       unsigned long long *__restrict__ counters /* [0]=nseeds [1]=sum plen */)
{ __shared__ unsigned wsum[MG_THREADS/32];
  __shared__ unsigned long long blockbase;

  unsigned i = blockIdx.x * MG_THREADS + threadIdx.x;
  unsigned cnt = 0, lowi = 0;
  int plen = 0;
  rec128 r1; r1.lo = r1.hi = 0;

  if (i < n1)
    { r1 = ld_rec(T1 + i);
      if (((r1.lo >> 47) & 1) == 0)                       // forward-strand entries only (:921-928)
        { unsigned p  = KREC_PREFIX24(r1.hi);
          unsigned lo = pstart2[p], hi = pstart2[p+1];
          if (lo < hi)
            { u64 s1 = KREC_SUFFIX56(r1);
              unsigned a = lo, b = hi;                    // lower bound of s1 in T2[lo,hi)
              while (a < b)
                { unsigned m = (a + b) >> 1;
                  if (suffix_of(T2,m) < s1) a = m+1; else b = m;
                }
              int ll = (a > lo) ? lcp56(s1,suffix_of(T2,a-1)) : -1;
              int lr = (a < hi) ? lcp56(s1,suffix_of(T2,a))   : -1;
              int m  = ll > lr ? ll : lr;
              plen = 12 + m;
              //  block of T2 entries sharing the first m suffix bases with s1: walk out from a
              unsigned lft = a, rgt = a;
              int sh = 56 - 2*m;                           // m <= 28 -> sh >= 0
              u64 key = (sh >= 56) ? 0 : (s1 >> sh);
              while (lft > lo && rgt - lft < (unsigned) freq)
                { u64 s = suffix_of(T2,lft-1);
                  if (((sh >= 56) ? 0 : (s >> sh)) != key) break;
                  lft -= 1;
                }
              while (rgt < hi && rgt - lft < (unsigned) freq)
                { u64 s = suffix_of(T2,rgt);
                  if (((sh >= 56) ? 0 : (s >> sh)) != key) break;
                  rgt += 1;
                }
              if (rgt - lft < (unsigned) freq)             // |R| < FREQ (:799-823)
                { cnt  = rgt - lft;
                  lowi = lft;
                }
            }
        }
    }

  //  block-level compaction
  int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  unsigned inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    { unsigned t = __shfl_up_sync(0xffffffffu,inc,o);
      if (lane >= o) inc += t;
    }
  if (lane == 31) wsum[wp] = inc;
  __syncthreads();
  unsigned pre = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < MG_THREADS/32; k++)
    { if (k < wp) pre += wsum[k];
      tot += wsum[k];
    }
  if (threadIdx.x == 0)
    blockbase = tot ? atomicAdd(&counters[0],(unsigned long long) tot) : 0ull;
  //  sum of plen (for the "ave len" statistic, FastGA.c:2485-2492)
  unsigned long long pl = (unsigned long long) cnt * plen;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    pl += __shfl_down_sync(0xffffffffu,pl,o);
  if (lane == 0 && pl) atomicAdd(&counters[1],pl);
  __syncthreads();
  if (cnt == 0) return;

  unsigned long long o = blockbase + pre + inc - cnt;
  if (o + cnt > capacity) return;                         // host sees counters[0] > capacity, retries
  unsigned icont = (unsigned) (r1.lo >> 32) & 0x7fff;
  long long ipost = (long long) (unsigned) r1.lo;
  for (unsigned k = 0; k < cnt; k++)
    { rec128 r2 = ld_rec(T2 + lowi + k);
      unsigned cs = (unsigned) (r2.lo >> 32) & 0xffff;
      st_rec(seeds + o + k,
             make_seed(L,cs >> 15,icont,cs & 0x7fff,ipost,(long long) (unsigned) r2.lo,plen));
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
  if (n1 > 0)
    { unsigned nb = (unsigned) ((n1 + MG_THREADS - 1) / MG_THREADS);
      cudaEvent_t ea, eb;
      cudaEventCreate(&ea); cudaEventCreate(&eb);
      cudaEventRecord(ea,st);
      adaptamer_merge_kernel<<<nb,MG_THREADS,0,st>>>((const rec128 *) d_T1,(unsigned) n1,
                                                     (const rec128 *) d_T2,d_pstart2,freq,L,
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
