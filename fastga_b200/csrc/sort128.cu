// LSD radix sort of 128-bit records on a byte range of the key, hand-written for sm_100a.
//
// Replaces the reference's two CPU sorts on the hot path:
//   msd_sort   (MSDsort.c:404)  -- k-mer records of the GIX build, key = 40-mer (80 bits)
//   rmsd_sort  (RSDsort.c:292)  -- adaptive-seed records, key = (jcont, band, anti, drem, lcp)
// Both are in-place American-flag MSD sorts on byte-packed records (MSDsort.c:211-360); on the
// device every record is widened to one 16-byte word so each pass is a perfectly coalesced
// stream (read 16 B, write 16 B per record).  One pass = one Onesweep kernel: a stable scatter
// that ranks records inside a 4096-record tile with warp match-any multi-split, finds the tile's
// bases by decoupled look-back, stages the tile in shared memory in digit order and writes digit
// runs back coalesced, while counting the next pass's histogram.  HBM-bound integer work: no
// tensor cores.
#include "common.cuh"

#define SORT_THREADS 512
#define SORT_ITEMS   8
#define SORT_TILE    (SORT_THREADS*SORT_ITEMS)
#define SORT_WARPS   (SORT_THREADS/32)

static __device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    { unsigned t = __shfl_up_sync(0xffffffffu,v,o);
      if (lane >= o) v += t;
    }
  return v;
}

/***********************************************************************************************
 *  Onesweep pass: one kernel per key byte reads every record once and writes it once.
 *   - the GLOBAL digit histogram of a pass is produced by the previous pass (each tile counts
 *     the next byte while it holds the records; the first pass has a small histogram kernel);
 *   - the tile's base inside each digit bucket comes from a decoupled look-back over the
 *     per-tile digit counts (status word = count | flag<<62; 1 = tile aggregate, 2 = inclusive
 *     prefix), tiles being handed out by an atomic ticket so every predecessor is running.
 **********************************************************************************************/

#define ST_AGG  (1ull << 62)
#define ST_INC  (2ull << 62)
#define ST_MASK ((1ull << 62) - 1)

__global__ void __launch_bounds__(SORT_THREADS)
sort_ghist_kernel(const rec128 *__restrict__ in, long long n, int byte, unsigned long long *__restrict__ ghist)
{ __shared__ unsigned h[256];
  int tid = threadIdx.x;
  if (tid < 256) h[tid] = 0;
  __syncthreads();
  const unsigned long long *half = reinterpret_cast<const unsigned long long *>(in) + (byte >= 8);
  int sh = 8*(byte & 7);
  for (long long tile0 = (long long) blockIdx.x * SORT_TILE; tile0 < n; tile0 += (long long) gridDim.x * SORT_TILE)
    {
#pragma unroll
      for (int it = 0; it < SORT_ITEMS; it++)
        { long long idx = tile0 + it*SORT_THREADS + tid;
          bool valid = idx < n;
          unsigned d = 0x100u | (tid & 31);
          if (valid) d = (unsigned) ((half[2*idx] >> sh) & 0xff);
          unsigned peers = __match_any_sync(0xffffffffu,d);
          if (valid && (tid & 31) == __ffs(peers)-1) atomicAdd(&h[d],__popc(peers));
        }
    }
  __syncthreads();
  if (tid < 256 && h[tid]) atomicAdd(&ghist[tid],(unsigned long long) h[tid]);
}

//  exclusive scan of a 256-bin global histogram -> bin bases; zeroes the histogram of the pass
//  after it and the tile ticket.
__global__ void sort_bins_kernel(const unsigned long long *__restrict__ ghist, unsigned long long *__restrict__ binbase,
                                 unsigned long long *__restrict__ nexthist, unsigned *__restrict__ ticket)
{ __shared__ unsigned long long ws[8];
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  unsigned long long v = ghist[tid], inc = v;
  for (int o = 1; o < 32; o <<= 1)
    { unsigned long long t = __shfl_up_sync(0xffffffffu,inc,o);
      if (lane >= o) inc += t;
    }
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  unsigned long long pre = 0;
  for (int i = 0; i < w; i++) pre += ws[i];
  binbase[tid] = pre + inc - v;
  nexthist[tid] = 0;
  if (tid == 0) *ticket = 0;
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_onesweep_kernel(const rec128 *__restrict__ in, rec128 *__restrict__ out, long long n, int byte,
                     int next_byte /* -1: none */, const unsigned long long *__restrict__ binbase,
                     unsigned long long *__restrict__ nexthist, unsigned long long *status /* [ntiles][256] */,
                     unsigned *__restrict__ ticket)
{ extern __shared__ __align__(16) unsigned char smem_raw[];
  rec128   *tile   = reinterpret_cast<rec128 *>(smem_raw);
  unsigned *wcount = reinterpret_cast<unsigned *>(tile + SORT_TILE);   // [SORT_WARPS][256]
  unsigned *bexcl  = wcount + SORT_WARPS*256;                           // [256]
  unsigned *nhist  = bexcl + 256;                                       // [256] next byte
  unsigned *wtot   = nhist + 256;                                       // [8]
  unsigned long long *gbase = reinterpret_cast<unsigned long long *>(wtot + 8);   // [256]
  __shared__ unsigned tile_s;
  __shared__ __align__(8) unsigned long long tbar;

  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0)
    { unsigned t = atomicAdd(ticket,1u);
      tile_s = t;
      //  the tile is 4096 consecutive 128-bit records: fetched by one TMA bulk copy (UBLKCP)
      //  into the staging buffer while the block zeroes its counters
      long long t0 = (long long) t * SORT_TILE, rm = n - t0;
      unsigned nb = (unsigned) (rm < SORT_TILE ? rm : SORT_TILE) * 16u;
      mbar_init(&tbar,1);
      tma_load_1d(tile,in + t0,nb,&tbar);
    }
  for (int i = tid; i < SORT_WARPS*256; i += SORT_THREADS) wcount[i] = 0;
  if (tid < 256) nhist[tid] = 0;
  __syncthreads();
  const unsigned tileid = tile_s;
  long long tile0 = (long long) tileid * SORT_TILE;
  long long rem = n - tile0;
  int cnt = rem < SORT_TILE ? (int) rem : SORT_TILE;

  rec128   r[SORT_ITEMS];
  unsigned rank[SORT_ITEMS];
  int base = w*(32*SORT_ITEMS);
  unsigned *myc = wcount + w*256;
  mbar_wait(&tbar,0);
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++)
    { int idx = base + it*32 + lane;
      bool valid = idx < cnt;
      unsigned d = 0x100u | lane;
      if (valid)
        { r[it] = ld_rec(tile + idx);
          d = rec_byte(r[it],byte);
        }
      unsigned peers = __match_any_sync(0xffffffffu,d);
      int leader = __ffs(peers)-1;
      unsigned b = 0;
      if (valid && lane == leader)
        { b = myc[d];
          myc[d] = b + __popc(peers);
        }
      b = __shfl_sync(0xffffffffu,b,leader);
      rank[it] = b + __popc(peers & lanemask_lt());
      if (next_byte >= 0)
        { unsigned d2 = valid ? rec_byte(r[it],next_byte) : (0x100u | lane);
          unsigned p2 = __match_any_sync(0xffffffffu,d2);
          if (valid && lane == __ffs(p2)-1) atomicAdd(&nhist[d2],__popc(p2));
        }
      __syncwarp();
    }
  __syncthreads();

  unsigned c = 0, inc = 0;
  if (tid < 256)
    { unsigned sum = 0;
#pragma unroll
      for (int ww = 0; ww < SORT_WARPS; ww++)
        { unsigned t = wcount[ww*256+tid];
          wcount[ww*256+tid] = sum;
          sum += t;
        }
      c = sum;
      //  publish the tile aggregate, then look back for the exclusive prefix of this digit
      volatile unsigned long long *stt = status;
      unsigned long long *mine = status + (unsigned long long) tileid*256 + tid;
      if (tileid == 0)
        atomicExch(mine,ST_INC | c);
      else
        atomicExch(mine,ST_AGG | c);
      unsigned long long excl = 0;
      for (long long t = (long long) tileid - 1; t >= 0; )
        { unsigned long long v = stt[(unsigned long long) t*256 + tid];
          if ((v >> 62) == 0) continue;                        // predecessor not there yet: spin
          excl += v & ST_MASK;
          if (v & ST_INC) break;
          t -= 1;
        }
      if (tileid != 0) atomicExch(mine,ST_INC | (excl + c));
      inc = warp_incl_scan(c,lane);
      if (lane == 31) wtot[w] = inc;
      gbase[tid] = binbase[tid] + excl;                        // minus bexcl below
      if (next_byte >= 0 && nhist[tid]) atomicAdd(&nexthist[tid],(unsigned long long) nhist[tid]);
    }
  __syncthreads();
  if (tid < 256)
    { unsigned pre = 0;
      for (int i = 0; i < w; i++) pre += wtot[i];
      unsigned ex = pre + inc - c;
      bexcl[tid] = ex;
      gbase[tid] -= ex;
    }
  __syncthreads();

#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++)
    { int idx = base + it*32 + lane;
      if (idx < cnt)
        { unsigned d = rec_byte(r[it],byte);
          st_rec(tile + (bexcl[d] + myc[d] + rank[it]),r[it]);
        }
    }
  __syncthreads();

  for (int p = tid; p < cnt; p += SORT_THREADS)
    { rec128 v = ld_rec(tile + p);
      unsigned d = rec_byte(v,byte);
      st_rec(out + (gbase[d] + p),v);
    }
}

static const size_t ONESWEEP_SMEM = SORT_TILE*sizeof(rec128) + (SORT_WARPS*256 + 256 + 256 + 8)*sizeof(unsigned)
                                    + 256*sizeof(unsigned long long);

extern "C" long long fgb_sort128_tmp_bytes(long long n)
{ long long ntiles = (n + SORT_TILE - 1) / SORT_TILE;
  if (ntiles < 1) ntiles = 1;
  return 256*ntiles*8 + (3*256 + 16)*8;          // look-back status + 2 histograms + bin bases + ticket
}

//  Sorts n records on key bytes [byte_lo,byte_hi) of the 128-bit little-endian value, stable.
//  d_a holds the input; d_b is a same-size scratch.  Returns via *result_in_b where the sorted
//  data landed (0 = d_a, 1 = d_b).  All pointers are device pointers.

extern "C" int fgb_sort128_device(void *d_a, void *d_b, long long n, int byte_lo, int byte_hi,
                                  void *d_tmp, long long tmp_bytes, int *result_in_b, void *stream)
{ cudaStream_t st = (cudaStream_t) stream;
  if (n < 0 || byte_lo < 0 || byte_hi > 16 || byte_lo > byte_hi) return FGB_ERR_ARG;
  if (n >= 0xffffffffll) return FGB_ERR_LIMIT;
  *result_in_b = 0;
  if (n <= 1 || byte_lo == byte_hi) return FGB_OK;
  if (tmp_bytes < fgb_sort128_tmp_bytes(n)) return FGB_ERR_ARG;

  int ntiles = (int) ((n + SORT_TILE - 1) / SORT_TILE);
  unsigned long long *status = (unsigned long long *) d_tmp;
  unsigned long long *hist[2] = { status + 256ull*ntiles, status + 256ull*ntiles + 256 };
  unsigned long long *binbase = status + 256ull*ntiles + 512;
  unsigned *ticket = (unsigned *) (binbase + 256);

  static bool attr_set = false;
  if (!attr_set)
    { CUDA_TRY(cudaFuncSetAttribute(sort_onesweep_kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int) ONESWEEP_SMEM));
      attr_set = true;
    }

  rec128 *src = (rec128 *) d_a, *dst = (rec128 *) d_b;
  CUDA_TRY(cudaMemsetAsync(hist[0],0,256*8,st));
  { int nb = ntiles < 1184 ? ntiles : 1184;
    sort_ghist_kernel<<<nb,SORT_THREADS,0,st>>>(src,n,byte_lo,hist[0]);
    fgb_count_launch(1);
  }
  int cur = 0;
  for (int b = byte_lo; b < byte_hi; b++)
    { sort_bins_kernel<<<1,256,0,st>>>(hist[cur],binbase,hist[cur^1],ticket);
      CUDA_TRY(cudaMemsetAsync(status,0,256ull*ntiles*8,st));
      sort_onesweep_kernel<<<ntiles,SORT_THREADS,ONESWEEP_SMEM,st>>>(src,dst,n,b,(b+1 < byte_hi) ? b+1 : -1,
                                                                   binbase,hist[cur^1],status,ticket);
      fgb_count_launch(2);
      cur ^= 1;
      rec128 *t = src; src = dst; dst = t;
      *result_in_b ^= 1;
    }
  CUDA_TRY(cudaGetLastError());
  return FGB_OK;
}

/***********************************************************************************************
 *  Generic exclusive scan of a u32 array (reduce / scan-of-sums / downsweep), used for stream
 *  compaction of syncmer posts, seeds and work lists.
 **********************************************************************************************/

#define SCAN_THREADS 1024
#define SCAN_ITEMS   8
#define SCAN_TILE    (SCAN_THREADS*SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const unsigned *__restrict__ data, long long n, unsigned long long *__restrict__ sums)
{ __shared__ unsigned long long ws[32];
  long long t0 = (long long) blockIdx.x * SCAN_TILE;
  unsigned long long s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++)
    { long long idx = t0 + i*SCAN_THREADS + threadIdx.x;
      if (idx < n) s += data[idx];
    }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu,s,o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32)
    { s = ws[threadIdx.x];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu,s,o);
      if (threadIdx.x == 0) sums[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
scan_sums_kernel(unsigned long long *__restrict__ sums, int nb, unsigned long long *__restrict__ total)
{ __shared__ unsigned long long ws[32];
  __shared__ unsigned long long carry_s;
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024)
    { int i = base + tid;
      unsigned long long v = (i < nb) ? sums[i] : 0, inc = v;
      for (int o = 1; o < 32; o <<= 1)
        { unsigned long long t = __shfl_up_sync(0xffffffffu,inc,o);
          if (lane >= o) inc += t;
        }
      if (lane == 31) ws[w] = inc;
      __syncthreads();
      if (w == 0)
        { unsigned long long s = ws[lane], si = s;
          for (int o = 1; o < 32; o <<= 1)
            { unsigned long long t = __shfl_up_sync(0xffffffffu,si,o);
              if (lane >= o) si += t;
            }
          ws[lane] = si - s;
        }
      __syncthreads();
      unsigned long long ex = carry_s + ws[w] + inc - v;
      if (i < nb) sums[i] = ex;
      __syncthreads();
      if (tid == 1023) carry_s = ex + v;
      __syncthreads();
    }
  if (tid == 0 && total != NULL) *total = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_down_kernel(unsigned *__restrict__ data, long long n, const unsigned long long *__restrict__ sums)
{ __shared__ unsigned ws[32];
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  long long t0 = (long long) blockIdx.x * SCAN_TILE + (long long) tid * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++)
    { v[i] = (t0 + i < n) ? data[t0+i] : 0;
      s += v[i];
    }
  unsigned inc = warp_incl_scan(s,lane);
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0)
    { unsigned x = ws[lane];
      unsigned xi = warp_incl_scan(x,lane);
      ws[lane] = xi - x;
    }
  __syncthreads();
  unsigned ex = (unsigned) sums[blockIdx.x] + ws[w] + inc - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++)
    { if (t0 + i < n) data[t0+i] = ex;
      ex += v[i];
    }
}

long long fgb_dev_scan_tmp_bytes(long long n)
{ long long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb < 1) nb = 1;
  return nb * 8 + 64;
}

//  In-place exclusive scan (values mod 2^32); *d_total (device, may be NULL) gets the 64-bit sum.

int fgb_dev_exclusive_scan_u32(unsigned *d_data, long long n, unsigned long long *d_total,
                               void *d_tmp, long long tmp_bytes, cudaStream_t st)
{ if (n <= 0)
    { if (d_total) CUDA_TRY(cudaMemsetAsync(d_total,0,8,st));
      return FGB_OK;
    }
  if (tmp_bytes < fgb_dev_scan_tmp_bytes(n)) return FGB_ERR_ARG;
  int nb = (int) ((n + SCAN_TILE - 1) / SCAN_TILE);
  unsigned long long *sums = (unsigned long long *) d_tmp;
  scan_reduce_kernel<<<nb,SCAN_THREADS,0,st>>>(d_data,n,sums);
  scan_sums_kernel<<<1,1024,0,st>>>(sums,nb,d_total);
  scan_down_kernel<<<nb,SCAN_THREADS,0,st>>>(d_data,n,sums);
  fgb_count_launch(3);
  CUDA_TRY(cudaGetLastError());
  return FGB_OK;
}
