// FP64 Gauss-Jordan inversion, with and without partial pivoting: the device side of
//   std::vector<double> matrix_inversion_FP64(std::vector<double>, int)
//       /root/reference/matrix_inv_solution/matrix_inversion_solution/matrix_inversion/matrix_inversion_FP64.cpp:13
//   std::vector<double> matrix_inversion_no_pivots(std::vector<double>, int)
//       .../matrix_inversion_no_pivots.cpp:10  (findCrr :41, fixRowKernel :60, copyCirColumn :50, fixColumnKernel :15)
// (the entry points next to the FP32 hot path.)
//
// Two schedules of the same arithmetic, both bit-identical to oracle/gj_oracle.c:gj_inplace_f64:
//   * blocked (default, second half of this file): 64-column panels, explicit row interchanges, per panel one
//     pivot-row recurrence and one trailing update -- 16 N^3 / 64 bytes of HBM traffic
//   * unblocked (MATINV_FLAG_UNBLOCKED): the schedule of gj_unblocked.cu in double precision, three launches per column,
//     the rank-1 update streams the whole matrix once per column (16 N^3 bytes, HBM-bound); kept as the cross-check
//
// The pivot candidate cannot ride in one 64-bit key as in FP32 (|x| alone is 63 bits), so partial results are
// (magnitude bits, row, value) triples compared lexicographically: larger magnitude, then lower row.  The NaN rules are
// those of common.cuh:gj_mag -- a NaN candidate never wins, a NaN incumbent (row r itself) is never displaced.
#include "common.cuh"
#include "kernels.h"

struct __align__(8) PivCand {
    u64 mag;      // bits of |x| (0 for a NaN that is not the incumbent, all ones for a NaN incumbent)
    double val;   // the signed entry
    int row;      // 0x7FFFFFFF = no candidate
    int pad;
};

__device__ __forceinline__ u64 gj_mag64(double x, bool incumbent) {
    const double a = fabs(x);
    return (a == a) ? (u64)__double_as_longlong(a) : (incumbent ? ~0ull : 0ull);
}
__device__ __forceinline__ bool cand_better(u64 m1, int r1, u64 m2, int r2) { return m1 > m2 || (m1 == m2 && r1 < r2); }

__device__ __forceinline__ void warp_best(u64 &mag, int &row, double &val) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 m2 = __shfl_xor_sync(0xffffffffu, mag, o);
        const int r2 = __shfl_xor_sync(0xffffffffu, row, o);
        const double v2 = __shfl_xor_sync(0xffffffffu, val, o);
        if (cand_better(m2, r2, mag, row)) { mag = m2; row = r2; val = v2; }
    }
}

// CTA-wide best of (mag,row,val); valid in every thread of warp 0 afterwards (sm: 8 entries each)
__device__ __forceinline__ void block_best(u64 &mag, int &row, double &val, u64 *smag, int *srow, double *sval) {
    warp_best(mag, row, val);
    const int lin = threadIdx.y * blockDim.x + threadIdx.x;   // 1-D blocks and (32, 8) blocks alike
    const int lane = lin & 31, warp = lin >> 5;
    if (lane == 0) { smag[warp] = mag; srow[warp] = row; sval[warp] = val; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x * blockDim.y) >> 5;
        mag = (lane < nw) ? smag[lane] : 0ull;
        row = (lane < nw) ? srow[lane] : 0x7FFFFFFF;
        val = (lane < nw) ? sval[lane] : 0.0;
        warp_best(mag, row, val);
    }
}

// (1) one partial per 256 rows of column `col`, rows >= row0
__global__ void __launch_bounds__(256) argmax_f64_kernel(const double *__restrict__ W, long long ld, int n, int col, int row0,
                                                         PivCand *__restrict__ part) {
    __shared__ u64 smag[8];
    __shared__ int srow[8];
    __shared__ double sval[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    u64 mag = 0;
    int row = 0x7FFFFFFF;
    double val = 0.0;
    if (i >= row0 && i < n) {
        val = W[(long long)i * ld + col];
        mag = gj_mag64(val, i == row0);
        row = i;
    }
    block_best(mag, row, val, smag, srow, sval);
    if (threadIdx.x == 0) {
        PivCand c;
        c.mag = mag; c.val = val; c.row = row; c.pad = 0;
        part[blockIdx.x] = c;
    }
}

// (1') no pivoting: the candidate is the diagonal entry itself (findCrr, matrix_inversion_no_pivots.cpp:41)
__global__ void diag_f64_kernel(const double *__restrict__ W, long long ld, int r, int col, PivCand *__restrict__ part) {
    PivCand c;
    c.val = W[(long long)r * ld + col];
    c.mag = gj_mag64(c.val, true);
    c.row = r;
    c.pad = 0;
    part[0] = c;
}

// (2) fused row swap + true division of the pivot row; thread x handles column x of the two rows and row x of the
// multiplier column.  Every CTA reduces the partials redundantly (no grid sync).
__global__ void __launch_bounds__(256) swap_normalize_f64_kernel(double *__restrict__ W, long long ld, int n, int r,
                                                                 const PivCand *__restrict__ part, int nparts,
                                                                 double *__restrict__ urow, double *__restrict__ ccol,
                                                                 int *__restrict__ piv, int *__restrict__ info) {
    __shared__ u64 smag[8];
    __shared__ int srow[8];
    __shared__ double sval[8];
    __shared__ int sp;
    __shared__ double sv;
    u64 mag = 0;
    int row = 0x7FFFFFFF;
    double val = 0.0;
    for (int g = threadIdx.x; g < nparts; g += 256) {
        const PivCand c = part[g];
        if (cand_better(c.mag, c.row, mag, row)) { mag = c.mag; row = c.row; val = c.val; }
    }
    block_best(mag, row, val, smag, srow, sval);
    if (threadIdx.x == 0) { sp = row; sv = val; }
    __syncthreads();
    const int p = sp;
    const double v = sv;   // carried with the candidate: W[p][r] itself is rewritten below
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x == 0) {
        piv[r] = p;
        if ((v == 0.0 || !isfinite(v)) && *info == 0) *info = r + 1;
    }
    if (x >= n) return;
    if (x != r && x != p) ccol[x] = W[(long long)x * ld + r];
    const double rr = W[(long long)r * ld + x];
    const double pp = W[(long long)p * ld + x];
    const double u = (x == r) ? 1.0 / v : pp / v;
    urow[x] = u;
    W[(long long)r * ld + x] = u;
    if (p != r) W[(long long)p * ld + x] = rr;
    if (x == r) {
        ccol[r] = 0.0;
        if (p != r) ccol[p] = rr;
    }
}

// (3) W[i][j] <- fma(-c_i, u_j, W[i][j]), column r <- fma(-c_i, u_r, +0); grid (ceil(n/64), ceil(n/8)), block (32, 8),
// two consecutive columns per thread (ld is even, so the 16-byte accesses are aligned).
__global__ void __launch_bounds__(256) rank1_update_f64_kernel(double *__restrict__ W, long long ld, int n, int r,
                                                               const double *__restrict__ urow,
                                                               const double *__restrict__ ccol) {
    const int i = blockIdx.y * 8 + threadIdx.y;
    const int j0 = (blockIdx.x * 32 + threadIdx.x) * 2;
    if (i >= n || j0 >= n || i == r) return;
    const double c = ccol[i];
    double *row = W + (long long)i * ld;
    if (j0 + 1 < n) {
        double2 a = *reinterpret_cast<double2 *>(row + j0);
        const double2 u = *reinterpret_cast<const double2 *>(urow + j0);
        a.x = (j0 == r) ? fma(-c, u.x, 0.0) : fma(-c, u.x, a.x);
        a.y = (j0 + 1 == r) ? fma(-c, u.y, 0.0) : fma(-c, u.y, a.y);
        *reinterpret_cast<double2 *>(row + j0) = a;
    } else {
        const double u = urow[j0];
        row[j0] = (j0 == r) ? fma(-c, u, 0.0) : fma(-c, u, row[j0]);
    }
}

__global__ void __launch_bounds__(256) load_f64_kernel(const double *__restrict__ A, int n, double *__restrict__ W, long long ld) {
    const long long i = blockIdx.x;
    const int j = blockIdx.y * 256 + threadIdx.x;
    if (j >= ld) return;
    W[i * ld + j] = (j < n) ? A[i * (long long)n + j] : 0.0;
}

// X[i][j] = W[i][colsrc[j]] (deferred column permutation) + isfinite scan; one CTA per row, the row staged in shared
// memory when it fits.
__global__ void __launch_bounds__(512) extract_f64_kernel(const double *__restrict__ W, long long ld, int n,
                                                          const int *__restrict__ colsrc, double *__restrict__ X,
                                                          int *__restrict__ info, int check, int staged) {
    extern __shared__ double srow_d[];
    const long long i = blockIdx.x;
    const double *wr = W + i * ld;
    double *xr = X + i * (long long)n;
    bool bad = false;
    if (staged) {
        for (int j = threadIdx.x; j < n; j += 512) srow_d[j] = wr[j];
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += 512) {
            const double v = srow_d[colsrc[j]];
            bad |= !isfinite(v);
            xr[j] = v;
        }
    } else {
        for (int j = threadIdx.x; j < n; j += 512) {
            const double v = wr[colsrc[j]];
            bad |= !isfinite(v);
            xr[j] = v;
        }
    }
    if (check && __syncthreads_or(bad) && threadIdx.x == 0) atomicCAS(info, 0, -1);
}

// ================================================================================================
// Blocked right-looking schedule (default).  Panels of F64_NB columns; inside a panel the three per-column kernels
// touch only the panel's columns (n x 64 doubles, L2 resident) and the two pivot rows; the multipliers of every step are
// kept, transposed, in CT (CT[t][i] = multiplier of row i at step t of the panel, following its row through later
// interchanges).  After the panel: one recurrence kernel brings the 64 pivot rows up to date on all other columns and
// leaves the snapshots U, one trailing-update kernel applies W[i][j] <- fma(-CT[t][i], U[t][j], W[i][j]), t ascending,
// to everything else.  Every element sees the FMA chain of the unblocked schedule, so the result is bit-identical
// (tests compare the two paths and the oracle); HBM traffic drops from 16 N^3 to 16 N^3 / 64 bytes.
#define F64_NB 64
#define F64_RPC 32   // rows per CTA of the single-launch panel step (one arg-max partial per CTA)

// (2b) pivot step t of a panel starting at column k0 (kw columns wide): reduce the partials, interchange rows r and p
// on ALL columns, normalise the pivot row on the panel's columns only, record the multiplier column.
__global__ void __launch_bounds__(256) panel_pivot_f64_kernel(double *__restrict__ W, long long ld, int n, int r, int k0, int kw,
                                                              int t, const PivCand *__restrict__ part, int nparts,
                                                              double *__restrict__ CT, double *__restrict__ upan,
                                                              double *__restrict__ pv, int *__restrict__ piv,
                                                              int *__restrict__ info) {
    __shared__ u64 smag[8];
    __shared__ int srow[8];
    __shared__ double sval[8];
    __shared__ int sp;
    __shared__ double sv;
    u64 mag = 0;
    int row = 0x7FFFFFFF;
    double val = 0.0;
    for (int g = threadIdx.x; g < nparts; g += 256) {
        const PivCand c = part[g];
        if (cand_better(c.mag, c.row, mag, row)) { mag = c.mag; row = c.row; val = c.val; }
    }
    block_best(mag, row, val, smag, srow, sval);
    if (threadIdx.x == 0) { sp = row; sv = val; }
    __syncthreads();
    const int p = sp;
    const double v = sv;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x == 0) {
        piv[r] = p;
        pv[t] = v;
        if ((v == 0.0 || !isfinite(v)) && *info == 0) *info = r + 1;
    }
    if (x < t && p != r) {   // multipliers of the panel's earlier steps follow their rows
        const double a = CT[(long long)x * ld + r], b = CT[(long long)x * ld + p];
        CT[(long long)x * ld + r] = b;
        CT[(long long)x * ld + p] = a;
    }
    if (x >= n) return;
    double *ct = CT + (long long)t * ld;
    if (x != r && x != p) ct[x] = W[(long long)x * ld + r];
    const double rr = W[(long long)r * ld + x];
    const double pp = W[(long long)p * ld + x];
    const bool in_panel = x >= k0 && x < k0 + kw;
    const double nr = in_panel ? ((x == r) ? 1.0 / v : pp / v) : pp;
    if (in_panel) upan[x - k0] = nr;
    W[(long long)r * ld + x] = nr;
    if (p != r) W[(long long)p * ld + x] = rr;
    if (x == r) {
        ct[r] = 0.0;
        if (p != r) ct[p] = rr;
    }
}

// (3b) rank-1 update restricted to the panel's columns: one warp per row, two columns per lane.  With `next_part` the
// kernel also leaves the arg-max partials of the NEXT column (r+1, rows > r) -- one per CTA of 8 rows -- so that the
// following pivot step needs no search launch of its own (the first column of a panel still does: the trailing
// update has just rewritten it).
__global__ void __launch_bounds__(256) panel_rank1_f64_kernel(double *__restrict__ W, long long ld, int n, int r, int k0, int kw,
                                                              int t, const double *__restrict__ CT,
                                                              const double *__restrict__ upan,
                                                              PivCand *__restrict__ next_part) {
    __shared__ u64 smag[8];
    __shared__ int srow[8];
    __shared__ double sval[8];
    const int i = blockIdx.x * 8 + threadIdx.y;
    const bool active = i < n && i != r;
    u64 mag = 0;
    int row = 0x7FFFFFFF;
    double val = 0.0;
    if (active) {
        const double c = CT[(long long)t * ld + i];
        double *wr = W + (long long)i * ld + k0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int jj = 2 * threadIdx.x + q;
            if (jj < kw) {
                const double u = upan[jj];
                const double a = (k0 + jj == r) ? fma(-c, u, 0.0) : fma(-c, u, wr[jj]);
                wr[jj] = a;
                if (next_part && jj == t + 1 && i > r) { val = a; mag = gj_mag64(a, i == r + 1); row = i; }
            }
        }
    }
    if (next_part) {   // uniform across the CTA
        block_best(mag, row, val, smag, srow, sval);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            PivCand c;
            c.mag = mag; c.val = val; c.row = row; c.pad = 0;
            next_part[blockIdx.x] = c;
        }
    }
}

// (2c) pivot step t of a panel in ONE launch.  The panel's columns ping-pong between two buffers (`in` -> `out`; the first
// step reads them from W, the last one writes them back), so no thread reads what another one writes and the pivot step
// (2b) and the panel update (3b) need no kernel boundary between them.  Every CTA reduces the partials and forms the
// normalised pivot row on the panel's columns itself; CTA 0 does the bookkeeping; the interchange of rows r and p on the
// columns outside the panel is spread over all CTAs; one warp per row updates the panel's columns and the CTA leaves the
// arg-max partial of the next column.
__global__ void __launch_bounds__(256) panel_step_f64_kernel(const double *__restrict__ in, long long ld_in,
                                                             double *__restrict__ out, long long ld_out,
                                                             double *__restrict__ W, long long ld, int n, int r, int k0, int kw,
                                                             int t, const PivCand *__restrict__ part, int nparts,
                                                             PivCand *__restrict__ next_part, double *__restrict__ CT,
                                                             double *__restrict__ pv, int *__restrict__ piv,
                                                             int *__restrict__ info) {
    __shared__ u64 smag[8];
    __shared__ int srow[8];
    __shared__ double sval[8];
    __shared__ int sp;
    __shared__ double sv;
    const int lane = threadIdx.x, wy = threadIdx.y, lin = wy * 32 + lane;
    u64 mag = 0;
    int row = 0x7FFFFFFF;
    double val = 0.0;
    for (int g = lin; g < nparts; g += 256) {
        const PivCand c = part[g];
        if (cand_better(c.mag, c.row, mag, row)) { mag = c.mag; row = c.row; val = c.val; }
    }
    block_best(mag, row, val, smag, srow, sval);
    if (lin == 0) { sp = row; sv = val; }
    __syncthreads();
    const int p = sp;
    const double v = sv;
    if (blockIdx.x == 0) {
        if (lin == 0) {
            piv[r] = p;
            pv[t] = v;
            if ((v == 0.0 || !isfinite(v)) && *info == 0) *info = r + 1;
        }
        if (lin < t && p != r) {   // multipliers of the panel's earlier steps follow their rows
            const double a = CT[(long long)lin * ld + r], b = CT[(long long)lin * ld + p];
            CT[(long long)lin * ld + r] = b;
            CT[(long long)lin * ld + p] = a;
        }
    }
    if (p != r) {   // rows r and p on the columns outside the panel
        for (int c = blockIdx.x * 256 + lin; c < n; c += gridDim.x * 256) {
            if (c >= k0 && c < k0 + kw) continue;
            const double a = W[(long long)r * ld + c], b = W[(long long)p * ld + c];
            W[(long long)r * ld + c] = b;
            W[(long long)p * ld + c] = a;
        }
    }
    // the panel's columns: F64_RPC rows per CTA, one warp per row at a time, two columns per lane
    mag = 0; row = 0x7FFFFFFF; val = 0.0;
    double u[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int jj = 2 * lane + q;
        u[q] = (jj < kw) ? ((jj == t) ? 1.0 / v : in[(long long)p * ld_in + jj] / v) : 0.0;
    }
    for (int k = 0; k < F64_RPC / 8; k++) {
        const int i = blockIdx.x * F64_RPC + wy + 8 * k;
        if (i >= n) break;
        const int src = (i == p) ? r : i;   // row p receives what row r held
        const double c = (i == r) ? 0.0 : in[(long long)src * ld_in + t];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int jj = 2 * lane + q;
            if (jj < kw) {
                double a;
                if (i == r) a = u[q];
                else a = (jj == t) ? fma(-c, u[q], 0.0) : fma(-c, u[q], in[(long long)src * ld_in + jj]);
                out[(long long)i * ld_out + jj] = a;
                if (next_part && jj == t + 1 && i > r) {
                    const u64 mq = gj_mag64(a, i == r + 1);
                    if (cand_better(mq, i, mag, row)) { mag = mq; row = i; val = a; }
                }
            }
        }
        if (lane == 0) CT[(long long)t * ld + i] = c;
    }
    if (next_part) {
        block_best(mag, row, val, smag, srow, sval);
        if (lin == 0) {
            PivCand c;
            c.mag = mag; c.val = val; c.row = row; c.pad = 0;
            next_part[blockIdx.x] = c;
        }
    }
}

// (4) the kw pivot rows (physically rows k0..k0+kw-1 after the interchanges) on every column outside the panel: per
// column, in step order, u = x[t] / v_t (snapshot -> U), then x[t2] <- fma(-c[t2][t], u, x[t2]) for the other pivot rows.
struct F64RowblockSmem {
    double cp[F64_NB][F64_NB];   // cp[t][t2] = multiplier of pivot row t2 at step t
    double pv[F64_NB];
    double x[F64_NB][128];
};

__global__ void __launch_bounds__(128) rowblock_f64_kernel(double *__restrict__ W, long long ld, int n, int k0, int kw,
                                                           const double *__restrict__ CT, const double *__restrict__ pv,
                                                           double *__restrict__ U) {
    extern __shared__ __align__(16) unsigned char f64_smem[];
    F64RowblockSmem &s = *reinterpret_cast<F64RowblockSmem *>(f64_smem);
    const int tid = threadIdx.x;
    const int j = blockIdx.x * 128 + tid;
    for (int e = tid; e < kw * kw; e += 128) {
        const int t = e / kw, t2 = e - t * kw;
        s.cp[t][t2] = CT[(long long)t * ld + k0 + t2];
    }
    if (tid < kw) s.pv[tid] = pv[tid];
    const bool active = j < n && !(j >= k0 && j < k0 + kw);
    if (active)
        for (int t = 0; t < kw; t++) s.x[t][tid] = W[(long long)(k0 + t) * ld + j];
    __syncthreads();
    if (!active) return;
    for (int t = 0; t < kw; t++) {
        const double u = s.x[t][tid] / s.pv[t];
        U[(long long)t * ld + j] = u;
        s.x[t][tid] = u;
        for (int t2 = 0; t2 < kw; t2++)
            if (t2 != t) s.x[t2][tid] = fma(-s.cp[t][t2], u, s.x[t2][tid]);
    }
    for (int t = 0; t < kw; t++) W[(long long)(k0 + t) * ld + j] = s.x[t][tid];
}

// (5) trailing update: 128 x 64 tiles, 256 threads x (8 x 4) doubles; the tile column of the panel is skipped (k0 is a
// multiple of 64), the pivot rows are computed along and not stored.  Accumulators are seeded from W and take the kw
// FMAs in step order.  The whole K extent (kw <= 64) of both operands sits in shared memory: 96 KB, two CTAs per SM.
__device__ __forceinline__ void f64_cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

struct F64GemmSmem {
    double a[F64_NB][128];   // a[t][ii] = CT[t][i0 + ii]
    double b[F64_NB][64];    // b[t][jj] = U[t][j0 + jj]
};

__global__ void __launch_bounds__(256, 2) trailing_f64_kernel(double *__restrict__ W, long long ld, int n, int k0, int kw,
                                                              const double *__restrict__ CT, const double *__restrict__ U) {
    extern __shared__ __align__(16) unsigned char f64_smem[];
    F64GemmSmem &s = *reinterpret_cast<F64GemmSmem *>(f64_smem);
    const int skip = k0 / 64;
    int tj = blockIdx.x;
    tj += (tj >= skip);
    const int i0 = blockIdx.y * 128, j0 = tj * 64;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const bool full = (i0 + 128 <= n) && (j0 + 64 <= n);
    if (full) {
        // both operands with 16-byte cp.async: every copy is in flight at once (a guarded load/store loop pays one
        // global-memory latency per iteration)
        for (int e = tid; e < kw * 64; e += 256) {
            const int t = e >> 6, c = (e & 63) * 2;
            f64_cp_async16(&s.a[t][c], CT + (long long)t * ld + i0 + c);
        }
        for (int e = tid; e < kw * 32; e += 256) {
            const int t = e >> 5, c = (e & 31) * 2;
            f64_cp_async16(&s.b[t][c], U + (long long)t * ld + j0 + c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        for (int e = tid; e < kw * 128; e += 256) {
            const int t = e >> 7, c = e & 127;
            s.a[t][c] = (i0 + c < n) ? CT[(long long)t * ld + i0 + c] : 0.0;
        }
        for (int e = tid; e < kw * 64; e += 256) {
            const int t = e >> 6, c = e & 63;
            s.b[t][c] = (j0 + c < n) ? U[(long long)t * ld + j0 + c] : 0.0;
        }
    }
    double acc[8][4];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int i = i0 + ty * 8 + q;
        if (full) {
            const double2 c0 = *reinterpret_cast<const double2 *>(W + (long long)i * ld + j0 + tx * 4);
            const double2 c1 = *reinterpret_cast<const double2 *>(W + (long long)i * ld + j0 + tx * 4 + 2);
            acc[q][0] = c0.x; acc[q][1] = c0.y; acc[q][2] = c1.x; acc[q][3] = c1.y;
        } else {
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const int j = j0 + tx * 4 + w;
                acc[q][w] = (i < n && j < n) ? W[(long long)i * ld + j] : 0.0;
            }
        }
    }
    if (full) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < kw; t++) {
        double a[8], b[4];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(&s.a[t][ty * 8 + q]);
            a[q] = v.x; a[q + 1] = v.y;
        }
#pragma unroll
        for (int w = 0; w < 4; w += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(&s.b[t][tx * 4 + w]);
            b[w] = v.x; b[w + 1] = v.y;
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
#pragma unroll
            for (int w = 0; w < 4; w++) acc[q][w] = fma(-a[q], b[w], acc[q][w]);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int i = i0 + ty * 8 + q;
        if (i >= k0 && i < k0 + kw) continue;   // pivot rows: brought up to date by the recurrence kernel
        if (full) {
            *reinterpret_cast<double2 *>(W + (long long)i * ld + j0 + tx * 4) = make_double2(acc[q][0], acc[q][1]);
            *reinterpret_cast<double2 *>(W + (long long)i * ld + j0 + tx * 4 + 2) = make_double2(acc[q][2], acc[q][3]);
        } else {
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const int j = j0 + tx * 4 + w;
                if (i < n && j < n) W[(long long)i * ld + j] = acc[q][w];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
void f64_workspace_free(F64Workspace &w) {
    cudaFree(w.W); cudaFree(w.urow); cudaFree(w.ccol); cudaFree(w.part); cudaFree(w.piv); cudaFree(w.colsrc); cudaFree(w.info);
    cudaFree(w.io); cudaFree(w.CT); cudaFree(w.U); cudaFree(w.pv); cudaFree(w.upan); cudaFree(w.P[0]); cudaFree(w.P[1]);
    w = F64Workspace();
}

cudaError_t f64_workspace_ensure(F64Workspace &w, int n, bool with_io) {
    if (w.n != n) {
        f64_workspace_free(w);
        const long long ld = ((long long)n + 1) & ~1ll;
        cudaError_t e;
        if ((e = cudaMalloc(&w.W, sizeof(double) * ld * n)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.urow, sizeof(double) * ld)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.ccol, sizeof(double) * ld)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.part, sizeof(PivCand) * 2 * (size_t)((n + 7) / 8 + 1))) != cudaSuccess) return e;   // partials, double-buffered
        if ((e = cudaMalloc(&w.piv, sizeof(int) * (size_t)n)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.colsrc, sizeof(int) * (size_t)n)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.info, sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.CT, sizeof(double) * ld * F64_NB)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.U, sizeof(double) * ld * F64_NB)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.pv, sizeof(double) * F64_NB)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.upan, sizeof(double) * F64_NB)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.P[0], sizeof(double) * (size_t)n * F64_NB)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.P[1], sizeof(double) * (size_t)n * F64_NB)) != cudaSuccess) return e;
        w.n = n;
        w.ld = ld;
    }
    if (with_io && !w.io) {
        cudaError_t e = cudaMalloc(&w.io, sizeof(double) * (size_t)n * n);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Blocked schedule: load, then per panel kw x (search, pivot step, panel rank-1) + pivot-row recurrence + trailing update,
// then column permutation + extraction.  prof_event brackets the trailing updates.
int f64_invert_blocked_async(F64Workspace &w, const double *A_dev, int n, double *X_dev, int nopivot, int check, cudaStream_t st,
                             cudaEvent_t (*prof_event)()) {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(extract_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(rowblock_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F64RowblockSmem));
        cudaFuncSetAttribute(trailing_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F64GemmSmem));
    }
    const long long ld = w.ld;
    int launches = 0;
    cudaMemsetAsync(w.info, 0, sizeof(int), st);
    load_f64_kernel<<<dim3(n, (unsigned)((ld + 255) / 256)), 256, 0, st>>>(A_dev, n, w.W, ld);
    launches++;
    const int nparts = (n + 255) / 256;
    const int ntile = (n + 63) / 64;
    for (int k0 = 0; k0 < n; k0 += F64_NB) {
        const int kw = (n - k0 < F64_NB) ? n - k0 : F64_NB;
        const int nrow8 = (n + 7) / 8, nstep = (n + F64_RPC - 1) / F64_RPC;
        for (int t = 0; t < kw; t++) {
            const int r = k0 + t;
            if (kw == 1) {   // a one-column panel cannot ping-pong (it would read and write W): the two in-place kernels
                if (nopivot) diag_f64_kernel<<<1, 1, 0, st>>>(w.W, ld, r, r, w.part);
                else argmax_f64_kernel<<<nparts, 256, 0, st>>>(w.W, ld, n, r, r, w.part);
                panel_pivot_f64_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.W, ld, n, r, k0, kw, t, w.part, nopivot ? 1 : nparts, w.CT,
                                                                        w.upan, w.pv, w.piv, w.info);
                panel_rank1_f64_kernel<<<nrow8, dim3(32, 8), 0, st>>>(w.W, ld, n, r, k0, kw, t, w.CT, w.upan, nullptr);
                launches += 3;
                continue;
            }
            const double *in = (t == 0) ? w.W + k0 : w.P[(t - 1) & 1];
            const long long ld_in = (t == 0) ? ld : F64_NB;
            double *out = (t == kw - 1) ? w.W + k0 : w.P[t & 1];
            const long long ld_out = (t == kw - 1) ? ld : F64_NB;
            // the search: left behind by the previous step except for the first column of a panel (and without pivoting).
            // The partials are read at the top of the kernel and rewritten at its end: two buffers, alternating.
            PivCand *pin = w.part + (size_t)(t & 1) * (nstep + 1), *pout = w.part + (size_t)((t + 1) & 1) * (nstep + 1);
            int np = nstep;
            if (nopivot) { diag_f64_kernel<<<1, 1, 0, st>>>(in, ld_in, r, t, pin); np = 1; launches++; }
            else if (t == 0) { argmax_f64_kernel<<<nparts, 256, 0, st>>>(w.W, ld, n, r, r, pin); np = nparts; launches++; }
            const bool fuse = !nopivot && t + 1 < kw;
            panel_step_f64_kernel<<<nstep, dim3(32, 8), 0, st>>>(in, ld_in, out, ld_out, w.W, ld, n, r, k0, kw, t, pin, np,
                                                                 fuse ? pout : nullptr, w.CT, w.pv, w.piv, w.info);
            launches++;
        }
        if (n > kw) {
            rowblock_f64_kernel<<<(n + 127) / 128, 128, sizeof(F64RowblockSmem), st>>>(w.W, ld, n, k0, kw, w.CT, w.pv, w.U);
            if (prof_event) cudaEventRecord(prof_event(), st);
            trailing_f64_kernel<<<dim3(ntile - 1, (n + 127) / 128), 256, sizeof(F64GemmSmem), st>>>(w.W, ld, n, k0, kw, w.CT, w.U);
            if (prof_event) cudaEventRecord(prof_event(), st);
            launches += 2;
        }
    }
    launch_colperm_build(w.piv, n, w.colsrc, st);
    const size_t bytes = (size_t)n * sizeof(double);
    const int staged = bytes <= 200 * 1024;
    extract_f64_kernel<<<n, 512, staged ? bytes : 0, st>>>(w.W, ld, n, w.colsrc, X_dev, w.info, check, staged);
    return launches + 2;
}

// Unblocked schedule (MATINV_FLAG_UNBLOCKED, cross-check of the blocked one):
// load + n x (search, swap + normalise, rank-1 update) + column permutation + extraction.  Returns the launch count.
// prof_event (may be NULL) hands out events that are recorded before and after every rank-1 update.
int f64_invert_async(F64Workspace &w, const double *A_dev, int n, double *X_dev, int nopivot, int check, cudaStream_t st,
                     cudaEvent_t (*prof_event)()) {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(extract_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const long long ld = w.ld;
    int launches = 0;
    cudaMemsetAsync(w.info, 0, sizeof(int), st);
    load_f64_kernel<<<dim3(n, (unsigned)((ld + 255) / 256)), 256, 0, st>>>(A_dev, n, w.W, ld);
    launches++;
    const int nparts = (n + 255) / 256;
    for (int r = 0; r < n; r++) {
        if (nopivot) diag_f64_kernel<<<1, 1, 0, st>>>(w.W, ld, r, r, w.part);
        else argmax_f64_kernel<<<nparts, 256, 0, st>>>(w.W, ld, n, r, r, w.part);
        swap_normalize_f64_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.W, ld, n, r, w.part, nopivot ? 1 : nparts, w.urow, w.ccol,
                                                                   w.piv, w.info);
        if (prof_event) cudaEventRecord(prof_event(), st);   // bench.py: the rank-1 update is the kernel the roofline is quoted on
        rank1_update_f64_kernel<<<dim3((n + 63) / 64, (n + 7) / 8), dim3(32, 8), 0, st>>>(w.W, ld, n, r, w.urow, w.ccol);
        if (prof_event) cudaEventRecord(prof_event(), st);
        launches += 3;
    }
    launch_colperm_build(w.piv, n, w.colsrc, st);
    const size_t bytes = (size_t)n * sizeof(double);
    const int staged = bytes <= 200 * 1024;
    extract_f64_kernel<<<n, 512, staged ? bytes : 0, st>>>(w.W, ld, n, w.colsrc, X_dev, w.info, check, staged);
    return launches + 2;
}
