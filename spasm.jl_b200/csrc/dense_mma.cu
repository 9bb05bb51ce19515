// dense_mma.cu — the one real dense contraction of the hot path on the 5th-gen tensor cores:
//     C (M x N) = [C -] A (M x K) . B^T (N x K)   mod p,     p < 2^16
// as tcgen05.mma.kind::i8 tiles (SASS: UTCIMMA) fed by TMA, int32 accumulators in TMEM, reduction
// mod p in the epilogue.  This is the trailing update / row transformation of the dense Schur tail
// (replaces FFLAS fgemm behind spasm_ffpack_rref, prototype src/SpaSM.jl:805).
//
// Limbs.  Residues in [0,p) are split as a = a1*256 + a0 with a0,a1 unsigned bytes, so
//     a*b = a1*b1 * 2^16 + (a1*b0 + a0*b1) * 2^8 + a0*b0.
// Four u8 x u8 MMAs per K-step feed THREE int32 accumulators (hi, mid, lo) in TMEM
// (3 x 128 columns of the 512): |lo|,|hi| <= 255^2 K and |mid| <= 2*255^2 K stay below 2^31 for
// K <= 16512, which covers every K this library produces (dense_block_size, default 1000).
// The epilogue recombines hi*2^16 + mid*2^8 + lo in 64 bits and Barrett-reduces once.
//
// Kernel shape.  One CTA per 128x128 output tile, 6 warps: warp 0 = TMA producer (4 boxes of
// 128 rows x 128 B per stage, SWIZZLE_128B), warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (tcgen05.ld 32x32b, one output row per thread).  3-stage smem ring of
// 64 KB stages, full/empty mbarriers, tcgen05.commit releases a stage / publishes the accumulator.
#include <cuda.h>

#include "dense.cuh"

namespace sb {

static constexpr int TILE = 128;   // output tile M = 128 rows
static constexpr int BK = 128;     // bytes (= int8 elements) of K per stage: one 128B swizzle row
static constexpr int MMA_THREADS = 192;
// Limb configurations.  L = 2 (p < 2^16): N = 128, 3 accumulators, 3 stages.  L = 3 (p < 2^24) and L = 4 (p < 2^32,
// the reference admits primes up to 4294967291, src/SpaSM.jl:74): N = 64 so that the 2L-1 int32 accumulators
// (one per power 2^(8s), s = i + j) still fit the 512 TMEM columns; L^2 u8 x u8 MMAs per K step; 2 stages.
// |acc_s| <= min(s+1, 2L-1-s) * 255^2 * K < 2^31 bounds K per launch (gemm_max_k).
template <int L>
struct LimbCfg {
  static constexpr int TN = (L == 2) ? 128 : 64;
  static constexpr int STAGES = (L == 2) ? 3 : 2;
  static constexpr int NACC = 2 * L - 1;
  static constexpr int STAGE_BYTES = L * TILE * BK + L * TN * BK;
  static constexpr int MAXK = (L == 2) ? 16384 : (L == 3) ? 10880 : 8192;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// K-major, SWIZZLE_128B operand tile: 8-row atoms of 1024 B (SBO = 1024), LBO unused (1), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::i8, D = s32, A = B = u8, both K-major, M = 128, N = TN
__host__ __device__ constexpr uint32_t idesc_for(int tn) {
  return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(tn >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}

// Tile order.  The CTAs in flight work on consecutive tile numbers, so the order decides what the L2 has to
// hold: row-major order re-streams ALL of B (N x K limbs, up to 900 MB in the bench) for every row of tiles
// (ncu, M=32768 N=16384 K=4096: 22 GB of DRAM reads for 2.5 GB of operands).  Tiles are therefore walked in
// column groups of RASTER_GN tiles: inside a group the B tiles (16 x 128 rows x K, 16 MB at K = 4096) stay in
// L2 while the rows of A stream past once.
static constexpr int RASTER_GN = 16;
template <int TN>
__device__ __forceinline__ void tile_origin(long long tile, int tiles_m, int tiles_n, int &m0, int &n0) {
  const long long per_group = (long long)RASTER_GN * tiles_m;
  const int grp = (int)(tile / per_group);
  const int within = (int)(tile - (long long)grp * per_group);
  const int gw = min(RASTER_GN, tiles_n - grp * RASTER_GN);
  m0 = (within / gw) * TILE;
  n0 = (grp * RASTER_GN + within % gw) * TN;
}

struct TensorMaps {
  CUtensorMap a[4], b[4];  // limb planes of A and B
};

struct __align__(8) MmaShared {
  uint64_t full[3];
  uint64_t empty[3];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
};
static constexpr int HALF = 64;                                 // columns of C handled per epilogue pass
static constexpr int STAGE_SP = HALF + 1;                       // padded row of the epilogue staging tile
static constexpr int STAGING_BYTES = TILE * STAGE_SP * 4;       // 128 x 65 words

// PERSISTENT: gridDim.x CTAs (one per SM) loop over the output tiles.  Per tile: the producer streams
// K through the smem ring (it runs ahead into the next tile while the epilogue is busy), the MMA
// thread accumulates the 2L-1 partial sums in TMEM, the epilogue warps pull the whole accumulator into registers
// (reduced mod p), hand TMEM back at once (tmem_empty) so the next tile's MMAs start, and only then do
// the read-modify-write of C through a padded staging tile, 64 columns at a time.
template <bool SUB, int L>
__global__ void __launch_bounds__(MMA_THREADS, 1)
k_gemm_i8limb(const __grid_constant__ TensorMaps maps, uint32_t *__restrict__ C, long long ldc, int M, int N, int nkb, int tiles_m, int tiles_n,
              const int *__restrict__ rowmap, Fp F) {
  using Cfg = LimbCfg<L>;
  constexpr int TN = Cfg::TN, STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES, NACC = Cfg::NACC;
  constexpr uint32_t IDESC = idesc_for(TN);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *tiles = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint32_t *stage = (uint32_t *)(tiles + STAGES * STAGE_BYTES);
  MmaShared *sh = (MmaShared *)((uint8_t *)stage + STAGING_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (long long)tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    mbar_init(&sh->tmem_full, 1);
    mbar_init(&sh->tmem_empty, 4);  // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int l = 0; l < L; l++) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[l]) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sh->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      long long kbg = 0;  // k-blocks issued so far (ring position)
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int m0, n0;
        tile_origin<TN>(tile, tiles_m, tiles_n, m0, n0);
        for (int kb = 0; kb < nkb; kb++, kbg++) {
          const int s = (int)(kbg % STAGES);
          const uint32_t ph = (uint32_t)((kbg / STAGES) & 1);
          mbar_wait(&sh->empty[s], ph ^ 1);
          uint8_t *st = tiles + s * STAGE_BYTES;
          mbar_expect_tx(&sh->full[s], STAGE_BYTES);
#pragma unroll
          for (int l = 0; l < L; l++) tma_load_2d(st + l * TILE * BK, &maps.a[l], kb * BK, m0, &sh->full[s]);
#pragma unroll
          for (int l = 0; l < L; l++) tma_load_2d(st + L * TILE * BK + l * TN * BK, &maps.b[l], kb * BK, n0, &sh->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      long long kbg = 0;
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        mbar_wait(&sh->tmem_empty, (it & 1) ^ 1);  // the epilogue has pulled the previous accumulator out
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < nkb; kb++, kbg++) {
          const int s = (int)(kbg % STAGES);
          const uint32_t ph = (uint32_t)((kbg / STAGES) & 1);
          mbar_wait(&sh->full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(tiles + s * STAGE_BYTES);
          uint64_t ad[L], bd[L];
#pragma unroll
          for (int l = 0; l < L; l++) ad[l] = make_desc(base + l * TILE * BK), bd[l] = make_desc(base + L * TILE * BK + l * TN * BK);
#pragma unroll
          for (int ks = 0; ks < BK / 32; ks++) {
            const uint64_t adv = (uint64_t)(ks * 32 >> 4);  // 32 bytes of K per instruction
            const uint32_t cont = (kb | ks) ? 1u : 0u;
#pragma unroll
            for (int j = 0; j < L; j++)
#pragma unroll
              for (int i = 0; i < L; i++) {
                constexpr int dummy = 0;
                (void)dummy;
                const int sidx = i + j;
                // the first product issued for power s (smallest j) starts the accumulator of this tile
                const int j_first = (sidx - (L - 1)) > 0 ? (sidx - (L - 1)) : 0;
                const uint32_t acc = (j == j_first) ? cont : 1u;
                umma_i8(tmem + sidx * TN, ad[i] + adv, bd[j] + adv, IDESC, acc);
              }
          }
          umma_commit(&sh->empty[s]);  // frees the stage when these MMAs have read it
        }
        umma_commit(&sh->tmem_full);
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32); each thread owns one output row in TMEM.
    // The C values of a 128 x 64 block are PREFETCHED into registers as 16 coalesced 16-byte
    // loads per thread before the accumulator is even ready, so no global-memory latency is exposed;
    // the reduced accumulator is transposed through the padded staging tile to meet them.
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int e = warp - 2;
    const int sub = lane >> 4, l16 = lane & 15;  // two rows per warp instruction, 16 lanes x 16 B each
    const bool vec_ok = ((ldc & 3) == 0) && ((((uintptr_t)C) & 15) == 0);
    constexpr int NH = TN / HALF;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      int m0, n0;
      tile_origin<TN>(tile, tiles_m, tiles_n, m0, n0);
#pragma unroll 1
      for (int h = 0; h < NH; h++) {
        const int nh = n0 + h * HALF;
        uint4 cv[16];
        // interior tiles: 16 unconditional 16-byte loads (all in flight together).  Row r of the product is row
        // rowmap[r] of C when a map is given (the dense tail only updates the columns that are still live).
        const bool interior = vec_ok && (m0 + TILE <= M) && (nh + HALF <= N);
        uint32_t *crow[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int row = m0 + (i * 4 + e) * 2 + sub;
          const long long grow = (rowmap != nullptr) ? (long long)(row < M ? rowmap[row] : 0) : (long long)row;
          crow[i] = C + grow * ldc + nh + l16 * 4;
        }
        if (SUB) {
          if (interior) {
#pragma unroll
            for (int i = 0; i < 16; i++) cv[i] = *(const uint4 *)crow[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
              const int row = m0 + (i * 4 + e) * 2 + sub, col = nh + l16 * 4;
              cv[i] = make_uint4(0, 0, 0, 0);
              if (row < M) {
                const uint32_t *p = crow[i];
                if (col + 0 < N) cv[i].x = p[0];
                if (col + 1 < N) cv[i].y = p[1];
                if (col + 2 < N) cv[i].z = p[2];
                if (col + 3 < N) cv[i].w = p[3];
              }
            }
          }
        }
        if (h == 0) {
          mbar_wait(&sh->tmem_full, it & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // staging free (previous block read out)
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
          uint32_t acc[NACC][16];
#pragma unroll
          for (int sidx = 0; sidx < NACC; sidx++) tmem_ld16(lane_base + sidx * TN + h * HALF + c0, acc[sidx]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; j++) {
            uint32_t r;
            if (L == 2) {
              const unsigned long long v = ((unsigned long long)acc[2][j] << 16) + ((unsigned long long)acc[1][j] << 8) + acc[0][j];
              r = red64(v, F);
            } else if (L == 3) {
              // Horner in 64 bits: (a4 2^16 + a3 2^8 + a2) mod p, then (. 2^16 + a1 2^8 + a0) mod p
              unsigned long long v = ((unsigned long long)acc[4][j] << 16) + ((unsigned long long)acc[3][j] << 8) + acc[2][j];
              v = ((unsigned long long)red64(v, F) << 16) + ((unsigned long long)acc[1][j] << 8) + acc[0][j];
              r = red64(v, F);
            } else {
              unsigned long long v = ((unsigned long long)acc[6][j] << 16) + ((unsigned long long)acc[5][j] << 8) + acc[4][j];
              v = ((unsigned long long)red64(v, F) << 24) + ((unsigned long long)acc[3][j] << 16) + ((unsigned long long)acc[2][j] << 8) + acc[1][j];
              v = ((unsigned long long)red64(v, F) << 8) + acc[0][j];
              r = red64(v, F);
            }
            stage[trow * STAGE_SP + c0 + j] = r;
          }
        }
        if (h == NH - 1) {  // the accumulator is out of TMEM: the next tile's MMAs may start
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sh->tmem_empty)) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (interior) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const uint32_t *sp = stage + ((i * 4 + e) * 2 + sub) * STAGE_SP + l16 * 4;
            uint4 r = make_uint4(sp[0], sp[1], sp[2], sp[3]);
            if (SUB) {
              r.x = addmod(cv[i].x, negmod(r.x, F), F);
              r.y = addmod(cv[i].y, negmod(r.y, F), F);
              r.z = addmod(cv[i].z, negmod(r.z, F), F);
              r.w = addmod(cv[i].w, negmod(r.w, F), F);
            }
            *(uint4 *)crow[i] = r;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const int rr = (i * 4 + e) * 2 + sub, row = m0 + rr, col = nh + l16 * 4;
            if (row < M) {
              const uint32_t *sp = stage + rr * STAGE_SP + l16 * 4;
              uint4 r = make_uint4(sp[0], sp[1], sp[2], sp[3]);
              if (SUB) {
                r.x = addmod(cv[i].x, negmod(r.x, F), F);
                r.y = addmod(cv[i].y, negmod(r.y, F), F);
                r.z = addmod(cv[i].z, negmod(r.z, F), F);
                r.w = addmod(cv[i].w, negmod(r.w, F), F);
              }
              uint32_t *p = crow[i];
              if (col + 0 < N) p[0] = r.x;
              if (col + 1 < N) p[1] = r.y;
              if (col + 2 < N) p[2] = r.z;
              if (col + 3 < N) p[3] = r.w;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------ limb split (u32 residues -> L K-major u8 planes, zero padded)
struct PlanePtrs {
  uint8_t *p[4];
};
template <int L>
__global__ void k_split_limbs(const uint32_t *__restrict__ in, long long ld, int rows, int K, PlanePtrs planes, int rows_pad, int Kpad,
                              const int *__restrict__ rowmap) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // one thread per 4 k
  const int kq = Kpad >> 2;
  if (idx >= (long long)rows_pad * kq) return;
  const int r = (int)(idx / kq), k = (int)(idx % kq) * 4;
  uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
  if (r < rows) {
    const uint32_t *src = in + (long long)(rowmap ? rowmap[r] : r) * ld + k;
    v0 = k + 0 < K ? src[0] : 0u, v1 = k + 1 < K ? src[1] : 0u, v2 = k + 2 < K ? src[2] : 0u, v3 = k + 3 < K ? src[3] : 0u;
  }
#pragma unroll
  for (int l = 0; l < L; l++) {
    const int sh = 8 * l;
    *(uchar4 *)(planes.p[l] + (long long)r * Kpad + k) = make_uchar4((v0 >> sh) & 255, (v1 >> sh) & 255, (v2 >> sh) & 255, (v3 >> sh) & 255);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || p == nullptr) throw Error("cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
static CUtensorMap make_map(uint8_t *base, int rows_pad, int Kpad, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows_pad};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad};
  cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// statistics for the roofline (bench.py): int8 MACs issued and device time spent in the MMA kernel
double g_mma_ms = 0;
double g_mma_macs = 0;       // modular MACs
double g_mma_int8_macs = 0;  // int8 MACs issued for them (x L^2: 4, 9 or 16)
long long g_mma_calls = 0;
static bool g_mma_disabled = false;
int g_gemm_cta_limit = 0;  // > 0: the persistent kernel uses at most this many CTAs (the second stream of the dense tail leaves SMs to the first)

// optional timing of the tcgen05 launches (bench.py / SPASM_B200_PROFILE): events are recorded around each launch
// and only read when the statistics are queried — no host synchronisation on the launch path
static bool g_mma_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_ev_pending, g_ev_free;
static void mma_collect_timings() {
  for (auto &pr : g_ev_pending) {
    if (cudaEventSynchronize(pr.second) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) g_mma_ms += ms;
    }
    g_ev_free.push_back(pr);
  }
  g_ev_pending.clear();
}

template <int L>
static void launch_limb_gemm(uint32_t *C, long long ldc, int M, int N, const uint32_t *A, long long lda, const uint32_t *B, long long ldb, int K,
                             bool subtract, const Fp &F, const int *rowmap) {
  using Cfg = LimbCfg<L>;
  constexpr int TN = Cfg::TN;
  cudaStream_t s = stream();
  const int Mp = (M + TILE - 1) / TILE * TILE, Np = (N + TN - 1) / TN * TN, Kp = (K + BK - 1) / BK * BK;
  DBuf<uint8_t> pa((size_t)L * Mp * Kp), pb((size_t)L * Np * Kp);
  PlanePtrs A_, B_;
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int l = 0; l < 4; l++) A_.p[l] = B_.p[l] = nullptr;
  for (int l = 0; l < L; l++) {
    A_.p[l] = pa.p + (size_t)l * Mp * Kp, B_.p[l] = pb.p + (size_t)l * Np * Kp;
    maps.a[l] = make_map(A_.p[l], Mp, Kp, TILE);
    maps.b[l] = make_map(B_.p[l], Np, Kp, TN);
  }
  k_split_limbs<L><<<cdiv((long long)Mp * (Kp >> 2), 256), 256, 0, s>>>(A, lda, M, K, A_, Mp, Kp, rowmap);
  k_split_limbs<L><<<cdiv((long long)Np * (Kp >> 2), 256), 256, 0, s>>>(B, ldb, N, K, B_, Np, Kp, nullptr);
  const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + STAGING_BYTES + 1024 + sizeof(MmaShared) + 64;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_gemm_i8limb<true, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_gemm_i8limb<false, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int tiles_m = Mp / TILE, tiles_n = Np / TN;
  const int grid = (int)std::min<long long>((long long)tiles_m * tiles_n, g_gemm_cta_limit > 0 ? std::min(g_gemm_cta_limit, sm_count()) : sm_count());
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (g_mma_timing) {
    if (g_ev_pending.size() >= 4096) mma_collect_timings();
    if (!g_ev_free.empty()) {
      ev = g_ev_free.back();
      g_ev_free.pop_back();
    } else {
      CK(cudaEventCreate(&ev.first));
      CK(cudaEventCreate(&ev.second));
    }
    CK(cudaEventRecord(ev.first, s));
  }
  if (subtract)
    k_gemm_i8limb<true, L><<<grid, MMA_THREADS, smem, s>>>(maps, C, ldc, M, N, Kp / BK, tiles_m, tiles_n, rowmap, F);
  else
    k_gemm_i8limb<false, L><<<grid, MMA_THREADS, smem, s>>>(maps, C, ldc, M, N, Kp / BK, tiles_m, tiles_n, rowmap, F);
  CK(cudaGetLastError());
  if (g_mma_timing) {
    CK(cudaEventRecord(ev.second, s));
    g_ev_pending.push_back(ev);
  }
  g_mma_macs += (double)M * N * K;  // algorithmic (unpadded) modular MACs
  g_mma_int8_macs += (double)M * N * K * L * L;
  g_mma_calls++;
}

int gemm_limbs(const Fp &F) { return F.p < (1u << 16) ? 2 : F.p < (1u << 24) ? 3 : 4; }
int gemm_max_k(const Fp &F) {
  const int L = gemm_limbs(F);
  return L == 2 ? LimbCfg<2>::MAXK : L == 3 ? LimbCfg<3>::MAXK : LimbCfg<4>::MAXK;
}

bool gemm_nt_mma(uint32_t *C, long long ldc, int M, int N, const uint32_t *A, long long lda, const uint32_t *B, long long ldb, int K,
                 bool subtract, const Fp &F, const int *rowmap) {
  if (g_mma_disabled) return false;
  static bool env_checked = false;
  if (!env_checked) {
    env_checked = true;
    if (getenv("SPASM_B200_NO_MMA")) {
      g_mma_disabled = true;
      return false;
    }
  }
  if (K < 64 || K > gemm_max_k(F) || (long long)M * N < 4LL * TILE * TILE) return false;
  switch (gemm_limbs(F)) {
    case 2: launch_limb_gemm<2>(C, ldc, M, N, A, lda, B, ldb, K, subtract, F, rowmap); break;
    case 3: launch_limb_gemm<3>(C, ldc, M, N, A, lda, B, ldb, K, subtract, F, rowmap); break;
    default: launch_limb_gemm<4>(C, ldc, M, N, A, lda, B, ldb, K, subtract, F, rowmap); break;
  }
  return true;
}

}  // namespace sb

// {ms in the tcgen05 kernel, modular MACs it computed, calls, kernels launched by the library}
extern "C" void spasm_b200_mma_timing(int on) { sb::g_mma_timing = on != 0; }
extern "C" void spasm_b200_mma_stats(double *out, int reset) {
  sb::mma_collect_timings();
  out[0] = sb::g_mma_ms, out[1] = sb::g_mma_macs, out[2] = (double)sb::g_mma_calls, out[3] = (double)sb::g_launches;
  if (reset) sb::g_mma_ms = sb::g_mma_macs = sb::g_mma_int8_macs = 0, sb::g_mma_calls = 0, sb::g_launches = 0;
}

// ------------------------------------------------------------------ measured tensor-pipe peak
// Back-to-back tcgen05.mma.kind::i8 M128 x N256 x K32 (SASS UTCIMMA) on every SM, operands resident in shared
// memory (no TMA, no epilogue): the denominator SURVEY.md section 8d asks for.  One CTA per SM, one issuing
// thread, two 256-column accumulators used alternately; `iters` groups of 8 instructions per CTA.
namespace sb {
static constexpr uint32_t IDESC_N256 = (2u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__global__ void __launch_bounds__(128, 1) k_utcimma_peak(int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *tiles = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) ((uint32_t *)tiles)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(tiles);
    const uint64_t a = make_desc(base), b = make_desc(base + 128 * 128);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const uint64_t adv = (uint64_t)(ks * 32 >> 4);
        const uint32_t acc = (it | ks) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(a + adv), "l"(b + adv), "r"(IDESC_N256), "r"(acc)
            : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + 256),
            "l"(a + adv), "l"(b + adv), "r"(IDESC_N256), "r"(acc)
            : "memory");
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}
}  // namespace sb

// returns the measured int8 tensor throughput in TOP/s (2 ops per int8 MAC), best of `reps`; < 0 on failure
extern "C" double spasm_b200_utcimma_peak(int iters, int reps) {
  using namespace sb;
  try {
    ApiCall api_scope_;
    if (iters <= 0) iters = 4096;
    if (reps <= 0) reps = 5;
    const size_t smem = (size_t)(128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_utcimma_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0;
    for (int r = 0; r < reps + 1; r++) {
      CK(cudaEventRecord(e0, stream()));
      k_utcimma_peak<<<sm_count(), 128, smem, stream()>>>(iters);
      CK(cudaGetLastError());
      CK(cudaEventRecord(e1, stream()));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * 128 * 256 * 32 * 8.0 * iters * sm_count();
      if (r > 0) best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
  } catch (const std::exception &e) {
    errf("[spasm_b200] utcimma peak measurement failed: %s\n", e.what());
    return -1;
  }
}

// test / bench hook: C = [C -] A . B^T mod p on host arrays of residues in [0,p).
// path 0: as the library would choose; 1: CUDA-core kernel only.  Returns 1 if the tcgen05 kernel ran.
extern "C" int spasm_b200_gemm_nt_host(long long prime, int M, int N, int K, const unsigned *A, const unsigned *B, unsigned *C, int subtract,
                                       int path, double *ms_out) {
  using namespace sb;
  try {
    ApiCall api_scope_;
    Fp F = make_field(prime);
    DBuf<uint32_t> dA((size_t)M * K), dB((size_t)N * K), dC((size_t)M * N);
    dA.upload(A, (size_t)M * K);
    dB.upload(B, (size_t)N * K);
    dC.upload(C, (size_t)M * N);
    const long long calls0 = g_mma_calls;
    const bool saved = g_mma_disabled;
    if (path == 1) g_mma_disabled = true;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, stream()));
    gemm_nt(dC.p, N, M, N, dA.p, K, dB.p, K, K, subtract != 0, F);
    CK(cudaEventRecord(e1, stream()));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_out) *ms_out = ms;
    g_mma_disabled = saved;
    dC.download(C, (size_t)M * N);
    sync();
    return g_mma_calls > calls0 ? 1 : 0;
  } catch (const std::exception &e) {
    errf("[spasm_b200] gemm test failed: %s\n", e.what());
    return -1;
  }
}
