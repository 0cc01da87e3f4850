// gemm_tf32x3.cu -- D[m][n] = sum_k A[m][k] * B[n][k] in FP32 accuracy on the 5th-generation tensor cores (sm_100a).
//
// Used by the fused track head (track_head.cu): the three contractions of ROIPool -> Linear(92659, 4) forward and
// backward are plain "NT" GEMMs once the operands are laid out K-major, and the operands are OUR OWN intermediate
// layouts (written by the layout kernels of track_head.cu), so unlike the reference's 38x63 NCHW maps they can have
// 16-byte-aligned row pitches: TMA applies.
//
//   operands   plain FP32 matrices.  The split  hi = tf32_rn(v) (in place),  lo = v - hi (second tile)  is made in shared
//              memory by the eight epilogue warps -- otherwise idle during the main loop -- while the previous k-block's
//              MMAs run; it is an elementwise pass over the tile, so the lo tile inherits the swizzled layout.  Three
//              kind::tf32 MMAs per k-step: hi*hi + hi*lo + lo*hi ("3xTF32"; the dropped lo*lo term is 2^-22 relative).
//   staging    TMA only, and only the raw values: half the bytes of pre-split hi/lo planes, which is what bounded the
//              first version of this kernel (22 us for 1.77 GFLOP: 103 MB of operand traffic through L2).  One 2-D box
//              {32 floats, rows} per operand and k-block lands as the K-major SWIZZLE_128B layout tcgen05.mma reads
//              (128-byte rows, 8-row atoms 1024 bytes apart); out-of-range rows / k are zero-filled by the TMA unit, so
//              ragged M, N and K need no padding in memory.
//   roles      warp 0: TMA producer (one elected lane); warp 1: MMA issuer (one elected lane); warp 2: TMEM allocation;
//              warps 4-11: lo-split of every stage, then the epilogue (tcgen05.ld 32 lanes x 16 columns at a time).
//              mbarrier ring: full (TMA landed) -> split (lo written) -> MMAs -> empty (tcgen05.commit).
//   tiles      M tile 128 (TMEM lanes), N tile BN <= 256 (TMEM columns), k-block 32; grid = (M tiles, N tiles, K splits).
//              Split-K partials go to disjoint slabs of the output (summed in a fixed order by the consumer:
//              deterministic, no atomics).
//   epilogue   ROW: out[(split * Mrows + m) * ldo + n]  (a thread owns a row: 16-byte stores);
//              COL: out[n * ldo + m]                      (a warp's 32 lanes are 32 consecutive m: coalesced).
#include <cuda.h>

#include "gemm_tf32x3.cuh"

namespace d2t {

namespace {

constexpr int GM = 128;      // M tile = UMMA M
constexpr int GBK = 32;      // k-block: 32 floats = one 128-byte swizzle row
constexpr int GTHREADS = 384;   // warps 0-2: TMA / MMA / TMEM allocation, 3: idle, 4-11: lo split + epilogue
constexpr int GSPLIT_WARPS = 8;

__device__ __forceinline__ uint32_t g_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            g_smem(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(g_smem(bar)) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_tma_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(g_smem(bar))
        : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 128-byte rows, 8-row atoms 1024 bytes apart (SBO); the start
// address may point 32*ks bytes into the row to select a k-step (same encoding as corr_umma_bwd.cu, verified there).
__device__ __forceinline__ uint64_t g_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void g_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void g_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g_smem(bar)) : "memory");
}

template <int BN>
struct GCfg {
    static constexpr int A_BYTES = GM * 128;   // one plane (raw = hi, or lo) of the A k-block
    static constexpr int B_BYTES = BN * 128;
    static constexpr int RAW_BYTES = A_BYTES + B_BYTES;          // what TMA delivers per k-block: [A raw | B raw]
    // two rings: raw tiles (TMA prefetch depth) and lo tiles (written by the split warps).  These GEMMs are short --
    // 7-9 k-blocks per CTA -- so the kernel is bound by latency, not bandwidth: a third raw stage hides the TMA round trip
    // (~1.2 us) behind the split + MMAs of the two k-blocks in flight (22 -> 13 us for the track-head forward GEMM).
    static constexpr int LO_STAGES = 2;
    static constexpr int RAW_STAGES = (226 * 1024 - LO_STAGES * RAW_BYTES) / RAW_BYTES >= 4 ? 4
                                      : (226 * 1024 - LO_STAGES * RAW_BYTES) / RAW_BYTES;
    static constexpr int SMEM_BYTES = (RAW_STAGES + LO_STAGES) * RAW_BYTES;
    static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M = 128");
    static_assert(RAW_STAGES >= 2, "stage ring");
};

template <int BN>
__global__ void __launch_bounds__(GTHREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ out,
                   GemmArgs a) {
    using Cfg = GCfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int RS = Cfg::RAW_STAGES, LS = Cfg::LO_STAGES;
    __shared__ __align__(8) uint64_t bar_full[RS], bar_rawfree[RS], bar_split[LS], bar_lofree[LS], bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * BN, split = blockIdx.z;
    // a following launch that carries the programmatic-serialization attribute (i.e. one the caller knows to be independent of
    // this GEMM, like the pad copy of the track-head backward) may start while this grid runs; ignored otherwise
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // k-blocks of this split: an even share, the first `rem` splits take one more
    const int per = a.kblocks / a.splits, rem = a.kblocks % a.splits;
    const int kbBeg = split * per + min(split, rem);
    const int kbCnt = per + (split < rem ? 1 : 0);

    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem(&tmem_base_s)),
                     "r"((uint32_t)Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < RS; ++s) {
            g_mbar_init(&bar_full[s], 1);
            g_mbar_init(&bar_rawfree[s], 1);
        }
        for (int s = 0; s < LS; ++s) {
            g_mbar_init(&bar_split[s], GSPLIT_WARPS);    // one arrival per split warp
            g_mbar_init(&bar_lofree[s], 1);
        }
        g_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smemBase = g_smem(smem);

    if (warp == 0) {
        // ================================ TMA producer =========================================================
        if (lane == 0) {
            for (int i = 0; i < kbCnt; ++i) {
                const int s = i % RS;
                const uint32_t round = (uint32_t)(i / RS);
                g_mbar_wait(&bar_rawfree[s], (round & 1u) ^ 1u);   // the MMAs that read this stage last round are done
                g_mbar_expect_tx(&bar_full[s], (uint32_t)Cfg::RAW_BYTES);
                const uint32_t st = smemBase + (uint32_t)s * Cfg::RAW_BYTES;
                const int k0 = (kbBeg + i) * GBK;
                g_tma_2d(st, &mapA, k0, m0, &bar_full[s]);
                g_tma_2d(st + Cfg::A_BYTES, &mapB, k0, n0, &bar_full[s]);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ===========================================================
        for (int i = 0; i < kbCnt; ++i) {
            const int r = i % RS, l = i % LS;
            g_mbar_wait(&bar_split[l], (uint32_t)(i / LS) & 1u);   // raw tiles landed AND their lo parts are written
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t aHi = smemBase + (uint32_t)r * Cfg::RAW_BYTES, bHi = aHi + Cfg::A_BYTES;
                const uint32_t aLo = smemBase + (uint32_t)(RS + l) * Cfg::RAW_BYTES, bLo = aLo + Cfg::A_BYTES;
#pragma unroll
                for (int ks = 0; ks < GBK / 8; ++ks) {
                    const uint32_t ko = ks * 32;   // 8 tf32 = 32 bytes inside the 128-byte row
                    g_mma(tmem_base, g_desc(aHi + ko), g_desc(bHi + ko), Cfg::IDESC, (i > 0 || ks > 0) ? 1u : 0u);
                    g_mma(tmem_base, g_desc(aHi + ko), g_desc(bLo + ko), Cfg::IDESC, 1u);
                    g_mma(tmem_base, g_desc(aLo + ko), g_desc(bHi + ko), Cfg::IDESC, 1u);
                }
                g_commit(&bar_rawfree[r]);                  // both rings are released once these MMAs have read them
                g_commit(&bar_lofree[l]);
                if (i == kbCnt - 1) g_commit(&bar_acc);     // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ================================ lo split, then epilogue (warps 4-11) =================================
        // The split is what paces a k-block (measured 1.4 us with four warps against 0.8 us of MMAs): eight warps, two
        // per scheduler.  In the epilogue warp w drains TMEM lane quarter w % 4 (the hardware's access rule), warps 4-7
        // the first half of the columns and warps 8-11 the second.
        const int quarter = warp & 3;
        const int half = (warp - 4) >> 2;
        {
            const int t = tid - 128;   // 0..255
            for (int i = 0; i < kbCnt; ++i) {
                const int r = i % RS, l = i % LS;
                g_mbar_wait(&bar_full[r], (uint32_t)(i / RS) & 1u);
                g_mbar_wait(&bar_lofree[l], ((uint32_t)(i / LS) & 1u) ^ 1u);   // the MMAs that read this lo tile are done
                float4* raw = reinterpret_cast<float4*>(smem + (size_t)r * Cfg::RAW_BYTES);
                float4* lo = reinterpret_cast<float4*>(smem + (size_t)(RS + l) * Cfg::RAW_BYTES);
                // hi = tf32_rn(v) replaces the raw word in place (the tensor core would otherwise TRUNCATE v, which leaves a
                // same-signed lo of up to 2^-10 |v| and a biased, twice larger error: measured 2.4e-6 against 9e-7 * sum|a||b|)
#pragma unroll 4
                for (int e = t; e < Cfg::RAW_BYTES / 16; e += 32 * GSPLIT_WARPS) {
                    const float4 v = raw[e];
                    const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                    raw[e] = h;
                    lo[e] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) g_mbar_arrive(&bar_split[l]);
            }
        }
        const int m = m0 + quarter * 32 + lane;
        if (kbCnt > 0) {
            g_mbar_wait(&bar_acc, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const bool mok = m < a.M;
        constexpr int QN = BN / 16, QH = (QN + 1) / 2;
#pragma unroll 1
        for (int q = half * QH; q < min(QN, (half + 1) * QH); ++q) {
            uint32_t r[16];
            if (kbCnt > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int x = 0; x < 16; ++x) r[x] = 0u;   // a split with no k-blocks still defines its slab
            }
            const int nb = n0 + q * 16;
            if (!mok || nb >= a.N) continue;
            if (a.epilogue == GEMM_EPI_ROW) {
                float* dst = out + ((size_t)split * a.slab_rows + m) * a.ldo + nb;
                if (nb + 16 <= a.N) {
#pragma unroll
                    for (int x = 0; x < 16; x += 4)
                        *reinterpret_cast<float4*>(dst + x) = make_float4(__uint_as_float(r[x]), __uint_as_float(r[x + 1]),
                                                                          __uint_as_float(r[x + 2]), __uint_as_float(r[x + 3]));
                } else {
#pragma unroll
                    for (int x = 0; x < 16; ++x)
                        if (nb + x < a.N) dst[x] = __uint_as_float(r[x]);
                }
            } else {
                float* dst = out + (size_t)split * a.slab_rows * a.ldo + (size_t)nb * a.ldo + m;
                if (a.col_rows > 0) {   // batch of images stacked along M: image b = m / col_rows owns its own output block
                    const int b = m / a.col_rows;
                    dst = out + (size_t)b * a.col_stride + (size_t)nb * a.ldo + (m - b * a.col_rows);
                }
#pragma unroll
                for (int x = 0; x < 16; ++x)
                    if (nb + x < a.N) dst[(size_t)x * a.ldo] = __uint_as_float(r[x]);
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS));
    }
}

// ---- tensor maps -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;   // resolved through the runtime: the library does not link libcuda
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x K fp32 matrix, K contiguous, row pitch `ld` floats (16-byte multiple); box = {32 floats, boxRows}, SWIZZLE_128B
static int make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, int boxRows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("gemm_tf32x3: cuTensorMapEncodeTiled is not available from this driver");
        return D2T_ERR_CUDA;
    }
    D2T_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 4) == 0 && ld >= K,
                "gemm_tf32x3: operand base / pitch must be 16-byte aligned (ld=%d, K=%d)", ld, K);
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)GBK, (cuuint32_t)boxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tf32x3: cuTensorMapEncodeTiled failed with %d (rows=%d K=%d ld=%d box=%d)", (int)r, rows, K, ld, boxRows);
        return D2T_ERR_CUDA;
    }
    return D2T_OK;
}

template <int BN>
static int launch(const GemmOperand& A, const GemmOperand& B, float* out, const GemmArgs& a, cudaStream_t st) {
    using Cfg = GCfg<BN>;
    CUtensorMap mA, mB;
    int rc;
    if ((rc = make_map(&mA, A.ptr, A.rows, a.K, A.ld, GM))) return rc;
    if ((rc = make_map(&mB, B.ptr, B.rows, a.K, B.ld, BN))) return rc;
    const size_t smem = (size_t)Cfg::SMEM_BYTES + 1024;
    D2T_SMEM_OPTIN(gemm_tf32x3_kernel<BN>, smem);
    dim3 grid(ceil_div(a.M, GM), ceil_div(a.N, BN), a.splits);
    gemm_tf32x3_kernel<BN><<<grid, GTHREADS, smem, st>>>(mA, mB, out, a);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace

int gemm_tf32x3(const GemmOperand& A, const GemmOperand& B, float* out, int M, int N, int K, int ldo, int epilogue, int splits,
                int slab_rows, int bn, cudaStream_t st, int col_rows, long long col_stride) {
    D2T_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, "gemm_tf32x3: bad shape M=%d N=%d K=%d splits=%d", M, N, K, splits);
    D2T_REQUIRE(A.rows >= M && B.rows >= N, "gemm_tf32x3: operand has fewer rows than the problem");
    D2T_REQUIRE(epilogue == GEMM_EPI_COL || (ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0),
                "gemm_tf32x3: row-major output needs a 16-byte aligned pitch");
    GemmArgs a;
    a.M = M; a.N = N; a.K = K; a.ldo = ldo; a.epilogue = epilogue; a.splits = splits; a.slab_rows = slab_rows;
    a.kblocks = ceil_div(K, GBK);
    a.col_rows = col_rows; a.col_stride = col_stride;
    D2T_REQUIRE(col_rows == 0 || (epilogue == GEMM_EPI_COL && splits == 1), "gemm_tf32x3: batched output needs the COL epilogue, no split-K");
    switch (bn) {
        case 208: return launch<208>(A, B, out, a, st);
        case 256: return launch<256>(A, B, out, a, st);
        case 64: return launch<64>(A, B, out, a, st);
        default: break;
    }
    set_error("gemm_tf32x3: unsupported N tile %d", bn);
    return D2T_ERR_BAD_ARG;
}

}  // namespace d2t
