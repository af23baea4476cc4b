// Column-sharded single inversion: per-rank primitives (one process per GPU).
//
// The n x n matrix is dealt to `world` ranks by column blocks of 128 (block J lives on rank J % world, as local
// block J / world); every rank holds all rows of its columns, so row interchanges, the row-block recurrence and
// the trailing update are local.  The only exchange is one message per block step, produced by the owner of the
// panel and broadcast by the host (torch.distributed / NCCL):
//
//     [ CmT: 128 x npad multipliers | pv: 128 pivot values | piv: 128 pivot rows | info | PanelState ]
//
// matinv_shard_factor writes that message IN PLACE (the panel kernels take the message's sub-buffers as their
// outputs), matinv_shard_apply consumes it on every rank.  This is north_star's "per-step pivot index and pivot
// row broadcast" at panel granularity; the reference itself is single-device
// (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:239-250).
// Same kernels, same FMA chains as the single-GPU path => the sharded result is bit-identical to it.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/matinv_shim.h"
#include "common.cuh"
#include "kernels.h"

struct matinv_shard {
    int n, npad, rank, world;
    int nblk;     // global number of 128-column blocks
    int nlocal;   // blocks held by this rank
    long long lcols;
    float *Wl, *U, *P[2];
    int *piv, *info;
};

namespace {
struct MsgLayout {
    size_t cmt, pv, piv, info, ps, total;
};
MsgLayout msg_layout(int npad) {
    MsgLayout m;
    m.cmt = 0;
    m.pv = (size_t)MATINV_NB * npad * sizeof(float);
    m.piv = m.pv + MATINV_NB * sizeof(float);
    m.info = m.piv + MATINV_NB * sizeof(int);
    m.ps = m.info + 16;
    m.total = m.ps + sizeof(PanelState);
    m.total = (m.total + 255) / 256 * 256;
    return m;
}
__global__ void merge_info_kernel(int *info, const int *msg_info) {
    if (*info == 0 && *msg_info != 0) *info = *msg_info;
}
}  // namespace

// defined in matinv_shim.cu
int shim_fail(int code, const char *fmt, ...);
int shim_device_count();

#define SCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return shim_fail(MATINV_E_CUDA, "%s -> %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

extern "C" {

long long matinv_shard_panel_bytes(int n) {
    if (n <= 0) return 0;
    const int npad = ((n + MATINV_NB - 1) / MATINV_NB) * MATINV_NB;
    return (long long)msg_layout(npad).total;
}

int matinv_shard_create(int n, int rank, int world, matinv_shard_t **out) {
    if (!out || n <= 0 || world <= 0 || rank < 0 || rank >= world) return shim_fail(MATINV_E_INVALID, "invalid argument");
    *out = nullptr;
    if (shim_device_count() == 0) return shim_fail(MATINV_E_NODEVICE, "no CUDA device");
    if (!subpanel_supported(n)) return shim_fail(MATINV_E_UNSUPPORTED, "n > 65536 is not supported by the sharded path");
    matinv_shard *s = new (std::nothrow) matinv_shard();
    if (!s) return shim_fail(MATINV_E_INVALID, "out of host memory");
    memset(s, 0, sizeof(*s));
    s->n = n; s->rank = rank; s->world = world;
    s->npad = ((n + MATINV_NB - 1) / MATINV_NB) * MATINV_NB;
    s->nblk = s->npad / MATINV_NB;
    s->nlocal = (s->nblk - rank + world - 1) / world;
    if (s->nlocal < 0) s->nlocal = 0;
    s->lcols = (long long)s->nlocal * MATINV_NB;
    const size_t N = (size_t)s->npad;
    const size_t lc = (size_t)(s->lcols > 0 ? s->lcols : MATINV_NB);
    cudaError_t e = cudaMalloc(&s->Wl, N * lc * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->U, MATINV_NB * lc * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->P[0], N * MATINV_NB * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->P[1], N * MATINV_NB * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->piv, N * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&s->info, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(s->Wl, 0, N * lc * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(s->U, 0, MATINV_NB * lc * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(s->P[0], 0, N * MATINV_NB * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(s->P[1], 0, N * MATINV_NB * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(s->piv, 0, N * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(s->info, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // null-stream memsets vs the caller's non-blocking streams
    if (e != cudaSuccess) {
        matinv_shard_destroy(s);
        return shim_fail(MATINV_E_CUDA, "shard allocation failed: %s", cudaGetErrorString(e));
    }
    *out = s;
    return MATINV_OK;
}

void matinv_shard_destroy(matinv_shard_t *s) {
    if (!s) return;
    cudaFree(s->Wl); cudaFree(s->U); cudaFree(s->P[0]); cudaFree(s->P[1]); cudaFree(s->piv); cudaFree(s->info);
    delete s;
}

float *matinv_shard_local(matinv_shard_t *s, long long *local_cols, long long *local_ld) {
    if (!s) return nullptr;
    if (local_cols) *local_cols = s->lcols;
    if (local_ld) *local_ld = s->lcols;
    return s->Wl;
}

int matinv_shard_generate(matinv_shard_t *s, unsigned long long seed, int kind, void *stream) {
    if (!s) return shim_fail(MATINV_E_INVALID, "invalid argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->lcols > 0) SCK(cudaMemsetAsync(s->Wl, 0, (size_t)s->npad * s->lcols * sizeof(float), st));
    SCK(cudaMemsetAsync(s->info, 0, sizeof(int), st));
    for (int jl = 0; jl < s->nlocal; jl++) {
        const int J = jl * s->world + s->rank;
        const int col0 = J * MATINV_NB;
        const int ncols = (s->n - col0 < MATINV_NB) ? s->n - col0 : MATINV_NB;
        if (ncols <= 0) continue;
        launch_generate(s->Wl + (size_t)jl * MATINV_NB, s->n, s->lcols, seed, kind, col0, ncols, st);
    }
    SCK(cudaGetLastError());
    return MATINV_OK;
}

// copy global column block J (n x 128, leading dimension ld) into / out of the local storage; src/dst may be host
// or device memory.  Only valid on the owner of J.
int matinv_shard_set_block(matinv_shard_t *s, int J, const float *src, long long ld, void *stream) {
    if (!s || !src || J < 0 || J >= s->nblk || J % s->world != s->rank) return shim_fail(MATINV_E_INVALID, "invalid argument");
    const int jl = J / s->world;
    const int ncols = (s->n - J * MATINV_NB < MATINV_NB) ? s->n - J * MATINV_NB : MATINV_NB;
    SCK(cudaMemcpy2DAsync(s->Wl + (size_t)jl * MATINV_NB, s->lcols * sizeof(float), src, ld * sizeof(float), ncols * sizeof(float),
                          s->n, cudaMemcpyDefault, (cudaStream_t)stream));
    SCK(cudaMemsetAsync(s->info, 0, sizeof(int), (cudaStream_t)stream));
    return MATINV_OK;
}
int matinv_shard_get_block(matinv_shard_t *s, int J, float *dst, long long ld, void *stream) {
    if (!s || !dst || J < 0 || J >= s->nblk || J % s->world != s->rank) return shim_fail(MATINV_E_INVALID, "invalid argument");
    const int jl = J / s->world;
    const int ncols = (s->n - J * MATINV_NB < MATINV_NB) ? s->n - J * MATINV_NB : MATINV_NB;
    SCK(cudaMemcpy2DAsync(dst, ld * sizeof(float), s->Wl + (size_t)jl * MATINV_NB, s->lcols * sizeof(float), ncols * sizeof(float),
                          s->n, cudaMemcpyDefault, (cudaStream_t)stream));
    return MATINV_OK;
}

int matinv_shard_factor(matinv_shard_t *s, int J, void *panel_dev, void *stream) {
    if (!s || !panel_dev || J < 0 || J >= s->nblk) return shim_fail(MATINV_E_INVALID, "invalid argument");
    if (J % s->world != s->rank) return shim_fail(MATINV_E_INVALID, "block %d is owned by rank %d", J, J % s->world);
    cudaStream_t st = (cudaStream_t)stream;
    const MsgLayout m = msg_layout(s->npad);
    char *msg = (char *)panel_dev;
    const int k0 = J * MATINV_NB;
    const int kb = (s->n - k0 < MATINV_NB) ? s->n - k0 : MATINV_NB;
    const int jl = J / s->world;
    SCK(cudaMemsetAsync(msg + m.pv, 0, m.total - m.pv, st));
    launch_panel_factor(s->Wl + (size_t)jl * MATINV_NB, s->lcols, s->n, k0, kb, (float *)(msg + m.cmt), s->npad,
                        (int *)(msg + m.piv) - k0, (float *)(msg + m.pv), (int *)(msg + m.info), (PanelState *)(msg + m.ps), s->P[0],
                        s->P[1], st);
    SCK(cudaGetLastError());
    return MATINV_OK;
}

// mode 0: every local column; mode 1: row-block step on every local column + trailing update of global block `block`
// only (must be local); mode 2: trailing update of every local column except `block`.  Modes 1 + 2 together equal mode 0 -- they exist so the owner of the NEXT panel can update that
// panel's columns first, factor it and broadcast it while the rest of the update is still running (look-ahead).
int matinv_shard_apply_ex(matinv_shard_t *s, int J, const void *panel_dev, void *stream, int mode, int block) {
    if (!s || !panel_dev || J < 0 || J >= s->nblk || mode < 0 || mode > 2) return shim_fail(MATINV_E_INVALID, "invalid argument");
    if (mode != 0 && (block < 0 || block >= s->nblk || block % s->world != s->rank)) return shim_fail(MATINV_E_INVALID, "block %d is not local", block);
    cudaStream_t st = (cudaStream_t)stream;
    const MsgLayout m = msg_layout(s->npad);
    const char *msg = (const char *)panel_dev;
    const float *CmT = (const float *)(msg + m.cmt), *pv = (const float *)(msg + m.pv);
    const PanelState *ps = (const PanelState *)(msg + m.ps);
    const int k0 = J * MATINV_NB;
    const int kb = (s->n - k0 < MATINV_NB) ? s->n - k0 : MATINV_NB;
    const int own_tile = (J % s->world == s->rank) ? J / s->world : -1;   // the panel's own tile, if it lives here
    const int nrt = s->npad / MATINV_NB;
    if (mode != 2) {
        SCK(cudaMemcpyAsync(s->piv + k0, msg + m.piv, kb * sizeof(int), cudaMemcpyDeviceToDevice, st));
        merge_info_kernel<<<1, 1, 0, st>>>(s->info, (const int *)(msg + m.info));
    }
    if (s->lcols > 0) {
        // row interchanges + recurrence always run on ALL local columns in the first call of a step (mode 0 / 1): the
        // look-ahead split only divides the trailing GEMM (block first, the rest while the next panel is factored)
        int skip = own_tile, skip_n = (own_tile >= 0) ? 1 : 0;
        if (mode != 2)
            launch_rowblock_ex(s->Wl, s->lcols, (int)s->lcols, k0, kb, skip, skip_n, CmT, s->npad, pv, ps, s->U, s->lcols, st);
        if (mode == 1) {
            const size_t off = (size_t)(block / s->world) * MATINV_NB;
            launch_trailing_gemm_ex(s->Wl + off, s->lcols, nrt, 1, J, -1, 0, kb, CmT, s->npad, s->U + off, s->lcols, st);
        } else {
            if (mode == 2) {
                const int bt = block / s->world;
                if (own_tile < 0) { skip = bt; skip_n = 1; }
                else if (bt == own_tile + 1) skip_n = 2;          // world == 1: panel tile and look-ahead tile are neighbours
                else return shim_fail(MATINV_E_INVALID, "look-ahead block must follow the panel");
            }
            launch_trailing_gemm_ex(s->Wl, s->lcols, nrt, s->nlocal, J, skip_n ? skip : -1, skip_n, kb, CmT, s->npad, s->U, s->lcols, st);
        }
    }
    SCK(cudaGetLastError());
    return MATINV_OK;
}

int matinv_shard_apply(matinv_shard_t *s, int J, const void *panel_dev, void *stream) {
    return matinv_shard_apply_ex(s, J, panel_dev, stream, 0, -1);
}

int matinv_shard_status(matinv_shard_t *s, int *info_host, int *piv_host, void *stream) {
    if (!s) return shim_fail(MATINV_E_INVALID, "invalid argument");
    cudaStream_t st = (cudaStream_t)stream;
    int info = 0;
    SCK(cudaMemcpyAsync(&info, s->info, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (piv_host) SCK(cudaMemcpyAsync(piv_host, s->piv, (size_t)s->n * sizeof(int), cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    if (info_host) *info_host = info;
    return info ? MATINV_SINGULAR : MATINV_OK;
}

}  // extern "C"
