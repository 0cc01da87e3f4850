// corr_common.cuh -- tile geometry and stream-K plan shared by the SIMT (corr_tile.cu) and tensor-core
// (corr_umma.cu) correlation kernels.
#pragma once
#include "common.cuh"

namespace d2t {

constexpr int kCorrThreads = 256;

template <int D>
struct FwdCfg {
    static constexpr int TD = 2 * D;              // live column / row displacements
    static constexpr int K1 = 2 * D + 1;          // output map side
    static constexpr int KK = K1 * K1;            // output map size
    static constexpr int QROWS = 128 / TD;        // query rows per tile
    static constexpr int QCOLS = 16;              // query cols per tile
    static constexpr int KROWS = QROWS + TD - 1;  // key rows per tile
    static constexpr int KCOLS = QCOLS + TD - 1;  // key cols per tile
    static constexpr int KP = (D == 8) ? 36 : 28; // key row pitch (floats): >= 16+TD, = 4 mod 8
    static constexpr int QP = 20;                 // query row pitch (floats): 16 used, = 4 mod 8
    static constexpr int KPATCH = KROWS * KP + 8;  // key rows >= 16 are shifted by 8 floats (bank-group skew)
    static constexpr int CH_FLOATS = QROWS * QP + KPATCH;  // staged floats per channel
    static constexpr int KV = (8 + TD) / 4;       // LDS.128 per thread per channel for keys
    static constexpr int KPASS = (KROWS + 7) / 8; // staging passes over key rows (8 rows x 32 lanes each)
    static constexpr int QPASS = (QROWS * QCOLS + kCorrThreads - 1) / kCorrThreads;
    static constexpr int TILE_FLOATS = QROWS * QCOLS * KK;     // one tile of output / one partial slot
    static_assert(D == 4 || D == 8, "tuned kernel covers d_max 4 and 8");
    static_assert(KP >= 16 + TD && KP % 4 == 0, "key pitch");
};

struct CorrPlan {
    int B, C, H, W;
    int tilesX, tilesY, T;  // tiles per image in x / y, total tiles
    int NI;                 // channel chunks per tile
    int G;                  // CTAs
    int ipc;                // (tile, chunk) iterations per CTA (backward: one contiguous range per CTA)
    // forward: `rounds` whole tiles per CTA (tile = round * G + cta, all CTAs walk the channels in step, so the halos
    // that neighbouring tiles share are fetched from L2, not HBM), then the `left` remaining tiles stream-K split
    int rounds, left, ipcL;
    int dbg;                // backward: chunks per channel group
    CorrOutStrides os;      // forward: output addressing (see common.cuh)
};

template <int D>
static inline int make_plan(int B, int C, int H, int W, int CK, CorrPlan* p) {
    using Cfg = FwdCfg<D>;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    p->B = B; p->C = C; p->H = H; p->W = W; p->dbg = 0;
    p->os.sb = (long long)H * W * Cfg::KK; p->os.sp = Cfg::KK; p->os.st = 1;
    p->tilesX = ceil_div(W, Cfg::QCOLS);
    p->tilesY = ceil_div(H, Cfg::QROWS);
    p->T = B * p->tilesX * p->tilesY;
    p->NI = ceil_div(C, CK);
    const long long total = (long long)p->T * p->NI;
    long long G = di.sm_count;
    if (G > total) G = total;
    p->G = (int)G;
    p->ipc = (int)((total + G - 1) / G);
    p->G = (int)((total + p->ipc - 1) / p->ipc);
    p->rounds = p->T / p->G;
    p->left = p->T - p->rounds * p->G;
    p->ipcL = (int)(((long long)p->left * p->NI + p->G - 1) / p->G);
    return 0;
}


// sums split-tile partial slots into `out` (defined in corr_tile.cu)
int corr_fwd_finalize8_launch(const float* partial, float* out, const CorrPlan& p, cudaStream_t st);

}  // namespace d2t
