// Single-call multi-GPU entries of the C-ABI: one process, one host thread per GPU, NCCL for the exchange.
//
//   matinv_invert_sharded_f32       one large matrix, 1-D block-cyclic column sharding (north_star:
//                                   "single large N uses 1-D block-cyclic column sharding ... broadcast over NVLink via NCCL")
//   matinv_invert_batched_f32_ngpu  batched small-n inversions split by matrix index, no communication
//
// The reference is single-device by construction (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:239-250:
// platforms[0] / devices[0], one in-order queue); these entries are what `matrix_inv_32` reaches when MATINV_NGPU > 1, so a
// C++ / Matlab caller gets more than one GPU behind the unchanged header.
//
// Sharded schedule = the per-rank primitives of gj_sharded.cu (matinv_shard_factor / _apply_ex: the same kernels and FMA
// chains as the single-GPU path, hence a bit-identical result) driven exactly like the torchrun host loop of
// gpu_matrix_inversion_b200/sharded.py:_factorize_lookahead, with ncclBroadcast on a high-priority side stream:
//   per 128-column block J:  owner(J+1) updates block J+1 first, factors it on the side stream and starts its broadcast
//                            while every rank (itself included) applies panel J to the remaining columns.
// After the last block the deferred column permutation X[:, j] = M[:, colsrc[j]] moves columns between ranks: a gather
// kernel packs, per destination rank, the columns it needs (row-major n x count), one grouped ncclSend / ncclRecv exchanges
// them, a scatter kernel places them.
//
// NCCL is loaded with dlopen at first use (libnccl.so.2: the copy already in the process if torch is loaded, else the
// system one), so libmatinv32.so keeps loading on machines without NCCL; there the multi-GPU entries return
// MATINV_E_UNSUPPORTED for ngpu > 1.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/matinv_shim.h"
#include "common.cuh"
#include "kernels.h"

int shim_fail(int code, const char *fmt, ...);
int shim_device_count();

namespace {

// ---------------------------------------------------------------------------------------------- NCCL, loaded lazily
struct Nccl {
    void *h = nullptr;
    bool tried = false;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
} g_nccl;

bool nccl_load() {
    if (g_nccl.tried) return g_nccl.h != nullptr;
    g_nccl.tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return false;
#define NCCL_SYM(field, name)                                   \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, #name);     \
    if (!g_nccl.field) { dlclose(h); return false; }
    NCCL_SYM(CommInitAll, ncclCommInitAll)
    NCCL_SYM(CommInitRankConfig, ncclCommInitRankConfig)
    NCCL_SYM(GetUniqueId, ncclGetUniqueId)
    NCCL_SYM(CommDestroy, ncclCommDestroy)
    NCCL_SYM(CommAbort, ncclCommAbort)
    NCCL_SYM(Broadcast, ncclBroadcast)
    NCCL_SYM(Send, ncclSend)
    NCCL_SYM(Recv, ncclRecv)
    NCCL_SYM(GroupStart, ncclGroupStart)
    NCCL_SYM(GroupEnd, ncclGroupEnd)
    NCCL_SYM(GetErrorString, ncclGetErrorString)
    NCCL_SYM(GetVersion, ncclGetVersion)
#undef NCCL_SYM
    g_nccl.h = h;
    return true;
}

// ---------------------------------------------------------------------------------------------- column exchange kernels
// sendbuf[i * cnt + k] = Wl[i * ld + cols[k]]   (row-major n x cnt, one thread per element)
__global__ void pack_columns_kernel(const float *__restrict__ Wl, long long ld, int n, const int *__restrict__ cols, int cnt,
                                    float *__restrict__ out) {
    const long long total = (long long)n * cnt;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / cnt;
        const int k = (int)(e - i * cnt);
        out[e] = Wl[i * ld + cols[k]];
    }
}
// Out[i * ld + cols[k]] = recvbuf[i * cnt + k]; *nonfinite is raised when an entry of the inverse is not finite (the scan the
// single-GPU extraction does, SURVEY A.2 -- on the device: scanning N^2 floats on one host thread costs a second at N = 32768)
__global__ void unpack_columns_kernel(float *__restrict__ Out, long long ld, int n, const int *__restrict__ cols, int cnt,
                                      const float *__restrict__ in, int *__restrict__ nonfinite) {
    const long long total = (long long)n * cnt;
    bool bad = false;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / cnt;
        const int k = (int)(e - i * cnt);
        const float v = in[e];
        bad |= !isfinite(v);
        Out[i * ld + cols[k]] = v;
    }
    if (bad) *nonfinite = 1;
}

// ---------------------------------------------------------------------------------------------- helpers
struct Barrier {
    std::mutex mu;
    std::condition_variable cv;
    int count, waiting = 0, gen = 0;
    explicit Barrier(int n) : count(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int my = gen;
        if (++waiting == count) {
            waiting = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != my; });
        }
    }
};

// colsrc with X[:, j] = M[:, colsrc[j]]: net effect of `for r = n-1..0: swap columns r, piv[r]` (SURVEY.md Appendix A.3;
// device twin csrc/gj_finish.cu:colperm_kernel, Python twin sharded.py:column_gather_list)
void column_gather_list(const int *piv, int n, std::vector<int> &idx) {
    idx.resize(n);
    for (int j = 0; j < n; j++) idx[j] = j;
    for (int r = n - 1; r >= 0; r--) {
        const int p = piv[r];
        if (p != r) std::swap(idx[r], idx[p]);
    }
}

// Two communicators per GPU.  `comm` (default configuration) carries the one grouped send/recv of the column permutation.
// `bcast` carries the per-block broadcasts and is limited to MATINV_MULTI_BCAST_CTAS (default 4) CTAs: a receiver's
// broadcast kernel is launched a whole block step ahead and spins until the owner has factored the panel, so every CTA it
// holds is an SM slot the trailing GEMM does not get for ~2.5 ms.  Measured on 8 B200 at N=65536 (factorisation + column
// exchange; MATINV_MULTI_TRACE splits busy / stalled time of the main stream): NCCL's default configuration 1493 ms (main
// stream busy 1474.6 ms, stalled 1.3 ms); maxCTAs = 1: 1738 ms (busy 1287 ms -- the interference is gone -- but stalled
// 425 ms, the broadcast is now too slow to hide); 2: 1375 ms; 4: 1314 ms = 7.88 x the 10.36 s of one GPU.
struct Comms {
    std::mutex mu;
    int ngpu = 0;
    std::vector<ncclComm_t> comm, bcast;
} g_comms;

int bcast_ctas() {
    const char *e = getenv("MATINV_MULTI_BCAST_CTAS");
    const int v = e ? atoi(e) : 4;
    return v;   // <= 0: NCCL's default
}

void destroy_comms() {
    for (ncclComm_t c : g_comms.bcast) if (c) g_nccl.CommDestroy(c);
    for (ncclComm_t c : g_comms.comm) if (c) g_nccl.CommDestroy(c);
    g_comms.bcast.clear();
    g_comms.comm.clear();
    g_comms.ngpu = 0;
}

int ensure_comms(int ngpu) {
    if (g_comms.ngpu == ngpu) return 0;
    if (!nccl_load()) return shim_fail(MATINV_E_UNSUPPORTED, "NCCL (libnccl.so.2) is not available: multi-GPU entries need it");
    if (g_comms.ngpu) destroy_comms();
    std::vector<int> devs(ngpu);
    for (int d = 0; d < ngpu; d++) devs[d] = d;
    g_comms.comm.assign(ngpu, nullptr);
    int cur = 0;
    cudaGetDevice(&cur);
    ncclResult_t r = g_nccl.CommInitAll(g_comms.comm.data(), ngpu, devs.data());
    if (r != ncclSuccess) {
        cudaSetDevice(cur);
        g_comms.comm.clear();
        return shim_fail(MATINV_E_CUDA, "ncclCommInitAll(%d) -> %s", ngpu, g_nccl.GetErrorString(r));
    }
    const int ctas = bcast_ctas();
    if (ctas > 0) {
        g_comms.bcast.assign(ngpu, nullptr);
        ncclUniqueId id;
        r = g_nccl.GetUniqueId(&id);
        if (r == ncclSuccess) r = g_nccl.GroupStart();
        for (int d = 0; d < ngpu && r == ncclSuccess; d++) {
            ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
            cfg.minCTAs = 1;
            cfg.maxCTAs = ctas;
            cudaSetDevice(d);
            r = g_nccl.CommInitRankConfig(&g_comms.bcast[d], ngpu, id, d, &cfg);
        }
        if (r == ncclSuccess) r = g_nccl.GroupEnd();
        if (r != ncclSuccess) {
            cudaSetDevice(cur);
            destroy_comms();
            return shim_fail(MATINV_E_CUDA, "ncclCommInitRankConfig(maxCTAs = %d) -> %s", ctas, g_nccl.GetErrorString(r));
        }
    }
    cudaSetDevice(cur);
    g_comms.ngpu = ngpu;
    return 0;
}

// Pinned bounce buffers of the host <-> shard transfers, two slots of one column block (n x 512 B) per device, cached across
// calls.  A rank's columns are every ngpu-th 512-byte segment of every row of the caller's matrix: copied straight from
// pageable memory (std::vector) the driver stages such strided 2-D copies on one thread per call -- measured 1.2 s of
// transfers around 0.25 s of compute for N = 32768 on 8 GPUs.  Here every rank's thread gathers / scatters the segments
// itself (8 threads in parallel) and the DMA of one block overlaps the gather of the next.
struct PinSlots {
    char *p = nullptr;
    size_t slot_bytes = 0;
} g_pin[64];

char *pin_slots(int dev, size_t slot_bytes) {
    PinSlots &ps = g_pin[dev & 63];
    if (ps.slot_bytes < slot_bytes) {
        if (ps.p) cudaFreeHost(ps.p);
        ps.p = nullptr;
        ps.slot_bytes = 0;
        if (cudaHostAlloc((void **)&ps.p, 2 * slot_bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ps.slot_bytes = slot_bytes;
    }
    return ps.p;
}

// rows [0, n) of a strided 2-D host copy (rb bytes per row), split over `helpers` threads: one thread moves 512-byte row
// segments at under 2 GB/s (TLB / cache misses on every row), and with few GPUs one thread per rank is far too little
void copy_rows_parallel(char *dst, size_t dst_pitch, const char *src, size_t src_pitch, size_t rb, int n, int helpers) {
    auto body = [&](int h) {
        const int lo = (int)((long long)n * h / helpers), hi = (int)((long long)n * (h + 1) / helpers);
        for (int i = lo; i < hi; i++) memcpy(dst + (size_t)i * dst_pitch, src + (size_t)i * src_pitch, rb);
    };
    if (helpers <= 1) { body(0); return; }
    std::vector<std::thread> th;
    for (int h = 1; h < helpers; h++) th.emplace_back(body, h);
    body(0);
    for (auto &t : th) t.join();
}
int transfer_helpers(int ngpu) {
    static int total = -1;
    if (total < 0) {
        const char *e = getenv("MATINV_MULTI_COPY_THREADS");
        total = e ? atoi(e) : 16;
        if (total < 1) total = 1;
    }
    const int k = total / (ngpu > 0 ? ngpu : 1);
    return k < 1 ? 1 : k;
}

struct RankState {
    matinv_shard_t *sh = nullptr;
    void *msg[2] = {nullptr, nullptr};
    cudaStream_t main = nullptr, side = nullptr;
    cudaEvent_t ev_top = nullptr, ev_ready = nullptr, ev_done = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_slot[2] = {nullptr, nullptr};
    float *sendbuf = nullptr, *recvbuf = nullptr, *Out = nullptr;
    int *cols_dev = nullptr;
    int rc = 0;
    char err[256] = "";
};

#define RCK(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) {                                                                              \
            snprintf(S.err, sizeof(S.err), "%s -> %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            S.rc = MATINV_E_CUDA;                                                                              \
            return;                                                                                            \
        }                                                                                                      \
    } while (0)
#define RNK(call)                                                                                              \
    do {                                                                                                       \
        ncclResult_t e__ = (call);                                                                             \
        if (e__ != ncclSuccess) {                                                                              \
            snprintf(S.err, sizeof(S.err), "%s -> %s (%s:%d)", #call, g_nccl.GetErrorString(e__), __FILE__, __LINE__); \
            S.rc = MATINV_E_CUDA;                                                                              \
            return;                                                                                            \
        }                                                                                                      \
    } while (0)
#define RMK(call)                                                                                              \
    do {                                                                                                       \
        const int rc__ = (call);                                                                               \
        if (rc__ < 0) {                                                                                        \
            snprintf(S.err, sizeof(S.err), "%s -> %d: %s", #call, rc__, matinv_last_error());                  \
            S.rc = rc__;                                                                                       \
            return;                                                                                            \
        }                                                                                                      \
    } while (0)

void rank_release(RankState &S) {
    if (S.sh) matinv_shard_destroy(S.sh);
    cudaFree(S.msg[0]); cudaFree(S.msg[1]); cudaFree(S.sendbuf); cudaFree(S.recvbuf); cudaFree(S.Out); cudaFree(S.cols_dev);
    if (S.ev_top) cudaEventDestroy(S.ev_top);
    if (S.ev_ready) cudaEventDestroy(S.ev_ready);
    if (S.ev_done) cudaEventDestroy(S.ev_done);
    if (S.ev_t0) cudaEventDestroy(S.ev_t0);
    if (S.ev_t1) cudaEventDestroy(S.ev_t1);
    if (S.ev_slot[0]) cudaEventDestroy(S.ev_slot[0]);
    if (S.ev_slot[1]) cudaEventDestroy(S.ev_slot[1]);
    if (S.main) cudaStreamDestroy(S.main);
    if (S.side) cudaStreamDestroy(S.side);
    S = RankState();
}

struct Shared {
    int n, ngpu, nblk;
    const float *A_host;
    float *X_host;
    int flags;
    unsigned long long gen_seed;   // A_host == NULL: synthetic workload generated on the devices (timing runs)
    int gen_kind;
    Barrier bar;
    std::vector<int> piv;          // rank 0's copy after the factorisation
    std::vector<int> colsrc;
    int info = 0;
    volatile int nonfinite = 0;    // an entry of the assembled inverse is not finite
    std::atomic<int> abort{0};     // a rank failed inside the block loop: the others stop issuing collectives
    bool exchange = true;          // apply the deferred column permutation across ranks (false: factorisation only)
    double compute_ms = 0.0;
    Shared(int g) : bar(g) {}
};

int owner_of(int J, int world) { return J % world; }
// position of global column c inside its owner's local storage
long long local_index(int c, int world) { return (long long)(c / MATINV_NB / world) * MATINV_NB + c % MATINV_NB; }

// Phase 1 (per rank): allocate, upload, factorise.  Phase 2 after a barrier: column exchange + download.
void rank_factor(int g, Shared &sh, RankState &S) {
    const int n = sh.n, G = sh.ngpu, nblk = sh.nblk;
    RCK(cudaSetDevice(g));
    int lo = 0, hi = 0;
    RCK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RCK(cudaStreamCreateWithFlags(&S.main, cudaStreamNonBlocking));
    RCK(cudaStreamCreateWithPriority(&S.side, cudaStreamNonBlocking, hi));
    RCK(cudaEventCreateWithFlags(&S.ev_top, cudaEventDisableTiming));
    RCK(cudaEventCreateWithFlags(&S.ev_ready, cudaEventDisableTiming));
    RCK(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
    RCK(cudaEventCreate(&S.ev_t0));
    RCK(cudaEventCreate(&S.ev_t1));
    RCK(cudaEventCreateWithFlags(&S.ev_slot[0], cudaEventDisableTiming));
    RCK(cudaEventCreateWithFlags(&S.ev_slot[1], cudaEventDisableTiming));
    RMK(matinv_shard_create(n, g, G, &S.sh));
    const long long mbytes = matinv_shard_panel_bytes(n);
    RCK(cudaMalloc(&S.msg[0], (size_t)mbytes));
    RCK(cudaMalloc(&S.msg[1], (size_t)mbytes));
    RCK(cudaMemset(S.msg[0], 0, (size_t)mbytes));
    RCK(cudaMemset(S.msg[1], 0, (size_t)mbytes));
    {   // exchange buffers (column permutation): at most every local column leaves and as many arrive
        long long lcols = 0, lld = 0;
        matinv_shard_local(S.sh, &lcols, &lld);
        const size_t cnt = (size_t)std::max<long long>(lcols, 1);
        RCK(cudaMalloc(&S.cols_dev, (2 * cnt + 1) * sizeof(int)));
        RCK(cudaMemset(S.cols_dev + 2 * cnt, 0, sizeof(int)));   // the non-finite flag lives behind the two column lists
        if (sh.exchange) {
            RCK(cudaMalloc(&S.sendbuf, cnt * n * sizeof(float)));
            RCK(cudaMalloc(&S.recvbuf, cnt * n * sizeof(float)));
            RCK(cudaMalloc(&S.Out, cnt * n * sizeof(float)));
        }
    }
    if (sh.A_host) {
        // gather block J (n rows x <= 128 columns) into a pinned slot, then one 2-D DMA into the shard; double buffered
        long long lcols = 0, lld = 0;
        float *Wl = matinv_shard_local(S.sh, &lcols, &lld);
        const size_t slot_bytes = (size_t)n * MATINV_NB * sizeof(float);
        char *pin = pin_slots(g, slot_bytes);
        if (!pin) { snprintf(S.err, sizeof(S.err), "cudaHostAlloc of the transfer slots failed"); S.rc = MATINV_E_CUDA; return; }
        int it = 0;
        for (int J = g; J < nblk; J += G, it++) {
            const int ncols = std::min(MATINV_NB, n - J * MATINV_NB);
            char *slot = pin + (size_t)(it & 1) * slot_bytes;
            if (it >= 2) RCK(cudaEventSynchronize(S.ev_slot[it & 1]));
            const size_t rb = (size_t)ncols * sizeof(float);
            const float *src = sh.A_host + (size_t)J * MATINV_NB;
            copy_rows_parallel(slot, rb, (const char *)src, (size_t)n * sizeof(float), rb, n, transfer_helpers(G));
            RCK(cudaMemcpy2DAsync(Wl + (size_t)(J / G) * MATINV_NB, (size_t)lld * sizeof(float), slot, rb, rb, (size_t)n, cudaMemcpyHostToDevice, S.main));
            RCK(cudaEventRecord(S.ev_slot[it & 1], S.main));
        }
    } else {
        RMK(matinv_shard_generate(S.sh, sh.gen_seed, sh.gen_kind, S.main));
    }
    RCK(cudaStreamSynchronize(S.main));
}

// MATINV_MULTI_TRACE=1: per rank, CUDA events on the main stream around the point where it waits for the side stream
// (next panel's message): "busy" = time the main stream spends in its own kernels, "stall" = time it waits for the
// message of the next panel.  Printed to stderr; diagnostic only.
bool multi_trace() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("MATINV_MULTI_TRACE");
        on = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return on == 1;
}

void rank_schedule(int g, Shared &sh, RankState &S) {
    const int G = sh.ngpu, nblk = sh.nblk;
    const bool trace = multi_trace();
    std::vector<cudaEvent_t> ea, eb;
    if (trace) {
        ea.resize(nblk); eb.resize(nblk);
        for (int J = 0; J < nblk; J++) { cudaEventCreate(&ea[J]); cudaEventCreate(&eb[J]); }
    }
    ncclComm_t comm = (G > 1) ? (g_comms.bcast.empty() ? g_comms.comm[g] : g_comms.bcast[g]) : nullptr;
    const size_t mbytes = (size_t)matinv_shard_panel_bytes(sh.n);
    RCK(cudaSetDevice(g));
    RCK(cudaEventRecord(S.ev_t0, S.main));
    if (owner_of(0, G) == g) RMK(matinv_shard_factor(S.sh, 0, S.msg[0], S.main));
    if (G > 1) RNK(g_nccl.Broadcast(S.msg[0], S.msg[0], mbytes, ncclChar, owner_of(0, G), comm, S.main));
    for (int J = 0; J < nblk; J++) {
        if (sh.abort.load(std::memory_order_relaxed)) {
            snprintf(S.err, sizeof(S.err), "aborted: another rank failed");
            S.rc = MATINV_E_CUDA;
            return;
        }
        void *msg = S.msg[J & 1];
        const int nxt = J + 1;
        if (nxt >= nblk) {
            RMK(matinv_shard_apply(S.sh, J, msg, S.main));
            break;
        }
        void *nmsg = S.msg[nxt & 1];
        const int own_n = owner_of(nxt, G);
        RCK(cudaEventRecord(S.ev_top, S.main));     // everything that read nmsg (the apply of panel J-1) is before this point
        if (own_n == g) {
            RMK(matinv_shard_apply_ex(S.sh, J, msg, S.main, 1, nxt));
            RCK(cudaEventRecord(S.ev_ready, S.main));
            RCK(cudaStreamWaitEvent(S.side, S.ev_ready, 0));
            RMK(matinv_shard_factor(S.sh, nxt, nmsg, S.side));
            if (G > 1) RNK(g_nccl.Broadcast(nmsg, nmsg, mbytes, ncclChar, own_n, comm, S.side));
            RMK(matinv_shard_apply_ex(S.sh, J, msg, S.main, 2, nxt));
        } else {
            RCK(cudaStreamWaitEvent(S.side, S.ev_top, 0));
            RNK(g_nccl.Broadcast(nmsg, nmsg, mbytes, ncclChar, own_n, comm, S.side));
            RMK(matinv_shard_apply(S.sh, J, msg, S.main));
        }
        RCK(cudaEventRecord(S.ev_done, S.side));
        if (trace) cudaEventRecord(ea[J], S.main);
        RCK(cudaStreamWaitEvent(S.main, S.ev_done, 0));
        if (trace) cudaEventRecord(eb[J], S.main);
    }
    RCK(cudaEventRecord(S.ev_t1, S.main));
    if (trace) {
        cudaStreamSynchronize(S.main);
        double busy = 0.0, stall = 0.0, stall_owner = 0.0;
        float ms = 0.f;
        for (int J = 0; J + 1 < nblk; J++) {
            if (cudaEventElapsedTime(&ms, ea[J], eb[J]) == cudaSuccess) { stall += ms; if (owner_of(J + 1, G) == g) stall_owner += ms; }
            if (J > 0 && cudaEventElapsedTime(&ms, eb[J - 1], ea[J]) == cudaSuccess) busy += ms;
        }
        float tot = 0.f;
        cudaEventElapsedTime(&tot, S.ev_t0, S.ev_t1);
        fprintf(stderr, "[matinv multi trace] rank %d/%d: total %.1f ms, main stream busy %.1f ms, waiting for the next panel's message %.1f ms (%.1f ms of it in steps where this rank factors)\n",
                g, G, tot, busy, stall, stall_owner);
        for (int J = 0; J < nblk; J++) { cudaEventDestroy(ea[J]); cudaEventDestroy(eb[J]); }
    }
}

void rank_status(int g, Shared &sh, RankState &S, std::vector<int> &piv, int &info) {
    RCK(cudaSetDevice(g));
    piv.resize(sh.n);
    const int rc = matinv_shard_status(S.sh, &info, piv.data(), S.main);
    if (rc < 0) {
        snprintf(S.err, sizeof(S.err), "matinv_shard_status -> %d: %s", rc, matinv_last_error());
        S.rc = rc;
    }
}

// Deferred column permutation across ranks + download of the local blocks of X (or, device-resident runs, nothing to
// download).  colsrc is shared (computed once by rank 0 from the pivot sequence every rank holds identically).
void rank_exchange(int g, Shared &sh, RankState &S) {
    const int n = sh.n, G = sh.ngpu, nblk = sh.nblk;
    ncclComm_t comm = (G > 1) ? g_comms.comm[g] : nullptr;
    RCK(cudaSetDevice(g));
    long long lcols = 0, lld = 0;
    float *Wl = matinv_shard_local(S.sh, &lcols, &lld);
    if (lcols == 0) return;
    // per peer: the local source columns I send (increasing destination column order) and the local slots I receive into
    std::vector<std::vector<int>> send_cols(G), recv_cols(G);
    for (int j = 0; j < n; j++) {
        const int src = sh.colsrc[j];
        const int sr = owner_of(src / MATINV_NB, G), dr = owner_of(j / MATINV_NB, G);
        if (sr == g) send_cols[dr].push_back((int)local_index(src, G));
        if (dr == g) recv_cols[sr].push_back((int)local_index(j, G));
    }
    size_t tot_send = 0, tot_recv = 0;
    std::vector<size_t> soff(G), roff(G);
    std::vector<int> flat;
    for (int p = 0; p < G; p++) { soff[p] = tot_send; tot_send += send_cols[p].size(); }
    for (int p = 0; p < G; p++) { roff[p] = tot_recv; tot_recv += recv_cols[p].size(); }
    for (int p = 0; p < G; p++) flat.insert(flat.end(), send_cols[p].begin(), send_cols[p].end());
    for (int p = 0; p < G; p++) flat.insert(flat.end(), recv_cols[p].begin(), recv_cols[p].end());
    RCK(cudaMemcpyAsync(S.cols_dev, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, S.main));
    for (int p = 0; p < G; p++) {
        const int cnt = (int)send_cols[p].size();
        if (!cnt) continue;
        const long long total = (long long)n * cnt;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        pack_columns_kernel<<<blocks, 256, 0, S.main>>>(Wl, lld, n, S.cols_dev + soff[p], cnt, S.sendbuf + soff[p] * n);
    }
    RCK(cudaGetLastError());
    if (G > 1) {
        RNK(g_nccl.GroupStart());
        for (int p = 0; p < G; p++) {
            if (p == g) continue;
            if (!send_cols[p].empty()) RNK(g_nccl.Send(S.sendbuf + soff[p] * n, send_cols[p].size() * (size_t)n, ncclFloat, p, comm, S.main));
            if (!recv_cols[p].empty()) RNK(g_nccl.Recv(S.recvbuf + roff[p] * n, recv_cols[p].size() * (size_t)n, ncclFloat, p, comm, S.main));
        }
        RNK(g_nccl.GroupEnd());
    }
    if (!send_cols[g].empty())
        RCK(cudaMemcpyAsync(S.recvbuf + roff[g] * n, S.sendbuf + soff[g] * n, send_cols[g].size() * (size_t)n * sizeof(float),
                            cudaMemcpyDeviceToDevice, S.main));
    for (int p = 0; p < G; p++) {
        const int cnt = (int)recv_cols[p].size();
        if (!cnt) continue;
        const long long total = (long long)n * cnt;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        unpack_columns_kernel<<<blocks, 256, 0, S.main>>>(S.Out, lcols, n, S.cols_dev + tot_send + roff[p], cnt, S.recvbuf + roff[p] * n,
                                                          S.cols_dev + 2 * (size_t)std::max<long long>(lcols, 1));
    }
    RCK(cudaGetLastError());
    RCK(cudaEventRecord(S.ev_t1, S.main));      // end of the device-resident window (factorisation + column exchange)
    if (sh.X_host) {
        // block by block through the pinned slots: the DMA of block b+1 runs while block b is scattered into the caller's rows
        const size_t slot_bytes = (size_t)n * MATINV_NB * sizeof(float);
        char *pin = pin_slots(g, slot_bytes);
        if (!pin) { snprintf(S.err, sizeof(S.err), "cudaHostAlloc of the transfer slots failed"); S.rc = MATINV_E_CUDA; return; }
        std::vector<int> mine;
        for (int J = g; J < nblk; J += G) mine.push_back(J);
        auto issue = [&](int it) -> cudaError_t {
            const int J = mine[it];
            const size_t rb = (size_t)std::min(MATINV_NB, n - J * MATINV_NB) * sizeof(float);
            cudaError_t e = cudaMemcpy2DAsync(pin + (size_t)(it & 1) * slot_bytes, rb, S.Out + (size_t)(J / G) * MATINV_NB,
                                              (size_t)lcols * sizeof(float), rb, (size_t)n, cudaMemcpyDeviceToHost, S.main);
            if (e == cudaSuccess) e = cudaEventRecord(S.ev_slot[it & 1], S.main);
            return e;
        };
        if (!mine.empty()) RCK(issue(0));
        for (int it = 0; it < (int)mine.size(); it++) {
            if (it + 1 < (int)mine.size()) RCK(issue(it + 1));
            RCK(cudaEventSynchronize(S.ev_slot[it & 1]));
            const int J = mine[it];
            const size_t rb = (size_t)std::min(MATINV_NB, n - J * MATINV_NB) * sizeof(float);
            const char *slot = pin + (size_t)(it & 1) * slot_bytes;
            float *dst = sh.X_host + (size_t)J * MATINV_NB;
            copy_rows_parallel((char *)dst, (size_t)n * sizeof(float), slot, rb, rb, n, transfer_helpers(G));
        }
    }
    int flag = 0;
    RCK(cudaMemcpyAsync(&flag, S.cols_dev + 2 * (size_t)std::max<long long>(lcols, 1), sizeof(int), cudaMemcpyDeviceToHost, S.main));
    RCK(cudaStreamSynchronize(S.main));
    if (flag) sh.nonfinite = 1;   // (ranks only ever write 1)
}

thread_local double g_last_sharded_ms = -1.0;

int run_sharded(const float *A_host, int n, float *X_host, int *piv_host, int ngpu, int flags, unsigned long long gen_seed,
                int gen_kind, bool exchange) {
    const int ndev = shim_device_count();
    if (ndev == 0) return shim_fail(MATINV_E_NODEVICE, "no CUDA device");
    if (ngpu <= 0) {
        const char *e = getenv("MATINV_NGPU");
        ngpu = e ? atoi(e) : ndev;
        if (ngpu <= 0) ngpu = ndev;
    }
    if (ngpu > ndev) return shim_fail(MATINV_E_INVALID, "ngpu = %d but only %d CUDA device(s) are visible", ngpu, ndev);
    const int nblk = (n + MATINV_NB - 1) / MATINV_NB;
    if (ngpu > nblk) ngpu = nblk;   // a rank without columns has nothing to do
    std::lock_guard<std::mutex> lk(g_comms.mu);   // one multi-GPU call at a time (the communicators are shared)
    if (ngpu > 1) {
        const int rc = ensure_comms(ngpu);
        if (rc) return rc;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    Shared sh(ngpu);
    sh.n = n; sh.ngpu = ngpu; sh.nblk = nblk; sh.A_host = A_host; sh.X_host = X_host; sh.flags = flags;
    sh.gen_seed = gen_seed; sh.gen_kind = gen_kind; sh.exchange = exchange;
    std::vector<RankState> st(ngpu);
    std::vector<std::vector<int>> pivs(ngpu);
    std::vector<int> infos(ngpu, 0);
    auto any_failed = [&]() {
        for (int g = 0; g < ngpu; g++) if (st[g].rc < 0) return true;
        return false;
    };
    auto body = [&](int g) {
        RankState &S = st[g];
        rank_factor(g, sh, S);
        sh.bar.wait();
        const bool ok1 = !any_failed();          // every rank sees the same verdict: all ranks wrote rc before the barrier
        if (ok1) rank_schedule(g, sh, S);
        if (ok1 && S.rc < 0 && !sh.abort.exchange(1) && ngpu > 1) {
            // This rank failed in the middle of the block loop: the collectives the others have already posted would wait
            // for it forever.  Aborting the communicators makes them return (the call fails as a whole and the
            // communicators are rebuilt by the next one).  The other ranks poll `abort` once per block step (~0.1-1 ms of
            // enqueue work): give them time to stop issuing collectives on the handles that are about to go away.
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
            for (ncclComm_t c : g_comms.bcast) if (c) g_nccl.CommAbort(c);
            for (ncclComm_t c : g_comms.comm) if (c) g_nccl.CommAbort(c);
            g_comms.bcast.clear();
            g_comms.comm.clear();
            g_comms.ngpu = 0;
        }
        if (ok1 && S.rc == 0) rank_status(g, sh, S, pivs[g], infos[g]);
        if (ok1 && S.rc == 0 && sh.abort.load()) {
            snprintf(S.err, sizeof(S.err), "aborted: another rank failed");
            S.rc = MATINV_E_CUDA;
        }
        sh.bar.wait();
        const bool ok2 = ok1 && !any_failed();
        if (ok2 && g == 0) {
            sh.info = infos[0];
            for (int r = 1; r < ngpu; r++)
                if (infos[r] != infos[0] || pivs[r] != pivs[0]) {
                    snprintf(S.err, sizeof(S.err), "ranks disagree on the pivot sequence / status word");
                    S.rc = MATINV_E_CUDA;
                }
            if (S.rc == 0 && sh.info == 0) column_gather_list(pivs[0].data(), n, sh.colsrc);
        }
        sh.bar.wait();
        if (ok2 && !any_failed() && sh.info == 0 && sh.exchange) rank_exchange(g, sh, S);
        if (S.rc == 0 && g == 0 && ok2) {
            float ms = 0.f;
            cudaSetDevice(0);
            if (cudaEventSynchronize(S.ev_t1) == cudaSuccess && cudaEventElapsedTime(&ms, S.ev_t0, S.ev_t1) == cudaSuccess) sh.compute_ms = ms;
        }
        sh.bar.wait();
    };
    std::vector<std::thread> th;
    for (int g = 1; g < ngpu; g++) th.emplace_back(body, g);
    body(0);
    for (auto &t : th) t.join();
    int rc = 0;
    char msg[320] = "";
    for (int g = 0; g < ngpu; g++)
        if (st[g].rc < 0 && rc == 0) {
            rc = st[g].rc;
            snprintf(msg, sizeof(msg), "GPU %d: %s", g, st[g].err);
        }
    if (rc == 0 && piv_host) memcpy(piv_host, pivs[0].data(), (size_t)n * sizeof(int));
    g_last_sharded_ms = (rc == 0) ? sh.compute_ms : -1.0;
    for (int g = 0; g < ngpu; g++) {
        cudaSetDevice(g);
        rank_release(st[g]);
    }
    cudaSetDevice(cur);
    if (rc < 0) {
        if (ngpu > 1 && g_comms.ngpu) {   // a failed rank may have left collectives half-issued: the communicators are not reusable
            for (ncclComm_t c : g_comms.bcast) if (c) g_nccl.CommAbort(c);
            for (ncclComm_t c : g_comms.comm) if (c) g_nccl.CommAbort(c);
            g_comms.bcast.clear();
            g_comms.comm.clear();
            g_comms.ngpu = 0;
        }
        return shim_fail(rc, "%s", msg);
    }
    if (sh.info != 0) {
        if (sh.info > 0) return shim_fail(MATINV_SINGULAR, "singular: zero or non-finite pivot at column %d", sh.info - 1);
        return shim_fail(MATINV_SINGULAR, "singular: non-finite entry in the inverse");
    }
    if (sh.nonfinite && !(flags & MATINV_FLAG_NOCHECK)) return shim_fail(MATINV_SINGULAR, "singular: non-finite entry in the inverse");
    return MATINV_OK;
}

}  // namespace

void multi_shutdown() {
    std::lock_guard<std::mutex> lk(g_comms.mu);
    for (PinSlots &ps : g_pin) {
        if (ps.p) cudaFreeHost(ps.p);
        ps = PinSlots();
    }
    if (g_comms.ngpu && g_nccl.h) destroy_comms();
    g_comms.bcast.clear();
    g_comms.comm.clear();
    g_comms.ngpu = 0;
}

extern "C" {

int matinv_nccl_version(void) {
    if (!nccl_load()) return 0;
    int v = 0;
    if (g_nccl.GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int matinv_invert_sharded_f32(const float *A_host, int n, float *X_host, int *piv_host, int ngpu, int nb, int flags) {
    if (n <= 0 || !A_host || !X_host) return shim_fail(MATINV_E_INVALID, "invalid argument");
    if (nb != 0 && nb != MATINV_NB) return shim_fail(MATINV_E_UNSUPPORTED, "column blocks are %d wide (nb = 0 selects the default)", MATINV_NB);
    if (flags & (MATINV_FLAG_TF32X3 | MATINV_FLAG_UNBLOCKED))
        return shim_fail(MATINV_E_UNSUPPORTED, "the column-sharded path runs the bit-exact blocked FP32 schedule only");
    return run_sharded(A_host, n, X_host, piv_host, ngpu, flags, 0ull, 0, true);   // (non-finite scan: unpack kernel)
}

int matinv_sharded_synthetic_f32(int n, unsigned long long seed, int kind, int ngpu, int *piv_host, double *compute_ms) {
    if (n <= 0) return shim_fail(MATINV_E_INVALID, "invalid argument");
    const int rc = run_sharded(nullptr, n, nullptr, piv_host, ngpu, 0, seed, kind, true);
    if (compute_ms) *compute_ms = g_last_sharded_ms;
    return rc;
}

int matinv_invert_batched_f32_ngpu(const float *A_host, int n, long long batch, float *X_host, int *info_host, int ngpu, int flags) {
    if (n <= 0 || n > 128 || batch < 0 || !A_host || !X_host) return shim_fail(MATINV_E_INVALID, "invalid argument (need 1 <= n <= 128)");
    if (batch == 0) return MATINV_OK;
    const int ndev = shim_device_count();
    if (ndev == 0) return shim_fail(MATINV_E_NODEVICE, "no CUDA device");
    if (ngpu <= 0) {
        const char *e = getenv("MATINV_NGPU");
        ngpu = e ? atoi(e) : ndev;
        if (ngpu <= 0) ngpu = ndev;
    }
    if (ngpu > ndev) return shim_fail(MATINV_E_INVALID, "ngpu = %d but only %d CUDA device(s) are visible", ngpu, ndev);
    if ((long long)ngpu > batch) ngpu = (int)batch;
    int cur = 0;
    cudaGetDevice(&cur);
    // contiguous index ranges [g * batch / G, (g+1) * batch / G): each GPU reads and writes only its slice
    std::vector<int> rcs(ngpu, 0);
    std::vector<std::string> errs(ngpu);
    auto body = [&](int g) {
        const long long b0 = batch * g / ngpu, b1 = batch * (g + 1) / ngpu, cnt = b1 - b0;
        if (cnt <= 0) return;
        cudaError_t e = cudaSetDevice(g);
        float *dA = nullptr;
        int *dI = nullptr;
        cudaStream_t st = nullptr;
        const size_t bytes = (size_t)cnt * n * n * sizeof(float);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&dA, bytes);
        if (e == cudaSuccess) e = cudaMalloc(&dI, (size_t)cnt * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpyAsync(dA, A_host + (size_t)b0 * n * n, bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = launch_batched(dA, n, cnt, dA, dI, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(X_host + (size_t)b0 * n * n, dA, bytes, cudaMemcpyDeviceToHost, st);
        std::vector<int> tmp;
        int *ih = info_host ? info_host + b0 : nullptr;
        if (!ih) { tmp.resize((size_t)cnt); ih = tmp.data(); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(ih, dI, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, st);
        const cudaError_t es = st ? cudaStreamSynchronize(st) : cudaSuccess;
        if (e == cudaSuccess) e = es;
        cudaFree(dA);
        cudaFree(dI);
        if (st) cudaStreamDestroy(st);
        if (e != cudaSuccess) {
            rcs[g] = MATINV_E_CUDA;
            errs[g] = cudaGetErrorString(e);
            return;
        }
        int any = 0;
        for (long long b = 0; b < cnt; b++) any |= (ih[b] != 0);
        rcs[g] = any ? MATINV_SINGULAR : MATINV_OK;
    };
    std::vector<std::thread> th;
    for (int g = 1; g < ngpu; g++) th.emplace_back(body, g);
    body(0);
    for (auto &t : th) t.join();
    cudaSetDevice(cur);
    int rc = MATINV_OK;
    for (int g = 0; g < ngpu; g++) {
        if (rcs[g] < 0) return shim_fail(rcs[g], "GPU %d: %s", g, errs[g].c_str());
        if (rcs[g] == MATINV_SINGULAR) rc = MATINV_SINGULAR;
    }
    return rc;
}

}  // extern "C"
