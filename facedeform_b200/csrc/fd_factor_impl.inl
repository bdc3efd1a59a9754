// fd_factor_impl.inl -- body of the LU kernels, included once per scalar type (see fd_factor.cu).
// REAL / REAL2: matrix scalar and its 2-vector; FD_LU_NS: namespace; FD_LU_NAME(x): exported launcher name.
namespace FD_LU_NS {

constexpr int NB = 32;            // block-column width
constexpr int PANEL_THREADS = 1024;
constexpr int PANEL_SMEM_MAX = 200 * 1024;

struct ArgMax {
    REAL v;
    int i;
};

__device__ __forceinline__ ArgMax argmax_combine(ArgMax a, ArgMax b)
{
    // larger magnitude wins; ties -> lower row index (deterministic, matches a serial first-max scan)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

// Unblocked LU with partial pivoting of the panel A[k0:n, k0:k0+nb].
// P points at the panel storage (global memory or the shared-memory copy) with leading dimension ldp.
template <bool IN_SMEM>
__global__ void __launch_bounds__(PANEL_THREADS) k_lu_panel(REAL* __restrict__ A, int lda, int n, int k0, int nb,
                                                            int* __restrict__ ipiv, int* __restrict__ flags,
                                                            double* __restrict__ pivstat)
{
    extern __shared__ REAL s_panel[];
    __shared__ ArgMax s_red[2][PANEL_THREADS / 32];
    __shared__ REAL s_pivrow[2][NB];
    const int m = n - k0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    REAL* G = A + (size_t)k0 * lda + k0; // panel origin in global memory
    const int ldp = IN_SMEM ? (m | 1) : lda;
    REAL* P = IN_SMEM ? s_panel : G;
    if (IN_SMEM) {
        for (int c = 0; c < nb; ++c)
            for (int r = tid; r < m; r += blockDim.x) P[(size_t)c * ldp + r] = G[(size_t)c * lda + r];
        __syncthreads();
    }
    double pmin = pivstat[0], pmax = pivstat[1];
    // Two block-wide barriers per column: every thread owns a fixed set of rows (r = tid mod blockDim), so the pivot
    // search of column j+1 only reads elements the same thread updated in column j.
    for (int j = 0; j < nb; ++j) {
        // (1) pivot search in column j, rows j..m-1
        ArgMax best = {-1.0, 0x7fffffff};
        const REAL* col = P + (size_t)j * ldp;
        for (int r = tid; r < m; r += blockDim.x) {
            if (r < j) continue;
            const REAL v = fabs(col[r]);
            if (v > best.v) best = {v, r}; // rows visited in increasing order per thread
        }
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        if (lane == 0) s_red[j & 1][warp] = best;
        __syncthreads();
        best = lane < nwarps ? s_red[j & 1][lane] : ArgMax{-1.0, 0x7fffffff}; // every warp reduces the partials itself
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        if (best.i >= m) best.i = j; // an all-NaN column: keep the diagonal, flagged singular below
        const int p = best.i;
        if (tid == 0) {
            ipiv[k0 + j] = k0 + p;
            if (!(best.v > 0.0) && flags[FD_FLAG_SINGULAR] == 0) flags[FD_FLAG_SINGULAR] = k0 + j + 1;
            pmin = fmin(pmin, (double)best.v);
            pmax = fmax(pmax, (double)best.v);
        }
        // (2) swap rows j and p inside the panel; keep the pivot row in shared memory
        if (tid < nb) {
            const REAL x = P[(size_t)tid * ldp + j], y = P[(size_t)tid * ldp + p];
            P[(size_t)tid * ldp + j] = y;
            P[(size_t)tid * ldp + p] = x;
            s_pivrow[j & 1][tid] = y;
        }
        __syncthreads();
        // (3) scale the column and rank-1 update the columns to its right (own rows only)
        const REAL piv = s_pivrow[j & 1][j];
        if (piv != 0.0) {
            const REAL inv = 1.0 / piv;
            for (int r = tid; r < m; r += blockDim.x) {
                if (r <= j) continue;
                const REAL l = P[(size_t)j * ldp + r] * inv;
                P[(size_t)j * ldp + r] = l;
#pragma unroll 4
                for (int c = j + 1; c < nb; ++c) P[(size_t)c * ldp + r] -= l * s_pivrow[j & 1][c];
            }
        }
    }
    __syncthreads();
    if (IN_SMEM) {
        for (int c = 0; c < nb; ++c)
            for (int r = tid; r < m; r += blockDim.x) G[(size_t)c * lda + r] = P[(size_t)c * ldp + r];
    }
    if (tid == 0) {
        pivstat[0] = pmin;
        pivstat[1] = pmax;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Register-resident, cluster-wide panel factorisation.
//
// The m x 32 panel lives in the register files of a thread-block cluster: CTA `cr` of the cluster owns panel rows
// [cr * 256 * RPT, (cr + 1) * 256 * RPT), thread t of it rows t, t + 256, ... (64 registers per row).  Per column:
//   local arg-max (registers -> warp REDUX -> CTA) -> candidates exchanged through distributed shared memory
//   -> cluster barrier -> every CTA picks the same pivot -> the owners of rows j and p publish them in their own
//   shared memory -> cluster barrier -> every CTA pulls both rows over DSMEM -> rank-1 update of its own rows.
// The row registers are shifted left by one column per step, fused into the update (a'[c-1] = a[c] - l * u[c]), so
// the loop over columns has static register indices and is not unrolled (an unrolled panel is ~200 KB of code that
// runs once, at instruction-fetch speed).  L and U entries go straight to global memory at the row's current
// position; the interchange inside the already finished columns is a fire-and-forget global swap by CTA 0.
// A cluster of 1 covers 256 * RPT rows (N = 256: one CTA); 16 CTAs x RPT = 3 cover 12288 rows.
// ---------------------------------------------------------------------------------------------------------------
struct Cand {
    REAL v;     // signed pivot candidate
    int i;        // panel row
    unsigned key; // high word of |v| (exponent + 20 mantissa bits): the magnitude the pivot search compares
};

// Pivot search on the 32-bit key with one REDUX per level instead of a 5-round shuffle tree on REALs: any entry
// within 2^-20 of the largest is as good a pivot for LU stability; ties go to the lowest lane, i.e. a fixed row order.
__device__ __forceinline__ unsigned mag_key(REAL v) { return fd_mag_key(v); }

__device__ __forceinline__ Cand cand_warp_pick(Cand c, bool valid)
{
    const unsigned key = valid ? c.key + 1u : 0u; // 0 = no candidate
    const unsigned best = __reduce_max_sync(0xffffffffu, key);
    const unsigned who = __ballot_sync(0xffffffffu, key == best);
    const int src = __ffs(who) - 1;
    Cand w;
    w.v = __shfl_sync(0xffffffffu, c.v, src);
    w.i = __shfl_sync(0xffffffffu, c.i, src);
    w.key = best ? best - 1u : 0u;
    if (best == 0u) w.i = 0x7fffffff;
    return w;
}

__device__ __forceinline__ REAL fast_rcp(REAL d) { return fd_fast_rcp(d); }

template <int RPT, bool CLUSTER>
__global__ void __launch_bounds__(256, 1) k_lu_panel_cluster(REAL* __restrict__ A, int lda, int n, int k0, int nb,
                                                             int* __restrict__ ipiv, int* __restrict__ flags,
                                                             double* __restrict__ pivstat, int* __restrict__ win)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int cr = CLUSTER ? (int)cluster.block_rank() : 0, CS = CLUSTER ? (int)cluster.num_blocks() : 1;
    __shared__ REAL s_fix[2 * NB][NB + 1]; // epilogue: window of the rows the interchanges touch
    __shared__ int s_rows[2 * NB];
    __shared__ int s_ib[NB];
    __shared__ int s_cnt;
    __shared__ int s_piv[NB];
    __shared__ Cand s_wred[2][8];      // per-warp candidates of this CTA
    __shared__ Cand s_cand[2][16];     // per-CTA candidates of the whole cluster (filled remotely)
    __shared__ __align__(16) REAL s_pub[2][2][NB]; // [parity][0: row j, 1: row p] published by the owner thread of this CTA
    __shared__ __align__(16) REAL s_row[2][2][NB]; // local copies pulled from the owners
    const int m = n - k0;
    const int rows_per_cta = 256 * RPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    REAL* G = A + (size_t)k0 * lda + k0;
    REAL a[RPT][NB];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int r = cr * rows_per_cta + tid + q * 256;
#pragma unroll
        for (int c = 0; c < NB; ++c) a[q][c] = (r < m && c < nb) ? G[(size_t)c * lda + r] : 0.0;
    }
    double pmin = INFINITY, pmax = 0.0;
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
        const int par = j & 1;
        // (1) local candidate
        Cand best = {0.0, 0x7fffffff, 0u};
        bool have = false;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = cr * rows_per_cta + tid + q * 256;
            const unsigned key = mag_key(a[q][0]);
            if (r >= j && r < m && (!have || key > best.key)) {
                best = Cand{a[q][0], r, key};
                have = true;
            }
        }
        best = cand_warp_pick(best, have);
        if (lane == 0) s_wred[par][warp] = best;
        __syncthreads();
        Cand win;
        if (CLUSTER) {
            if (warp == 0) {
                Cand c = lane < 8 ? s_wred[par][lane] : Cand{0.0, 0x7fffffff, 0u};
                c = cand_warp_pick(c, lane < 8 && c.i != 0x7fffffff);
                // (2) hand the CTA candidate to every CTA of the cluster
                if (lane < CS) *cluster.map_shared_rank(&s_cand[par][cr], lane) = c;
            }
            cluster.sync();
            win = lane < CS ? s_cand[par][lane] : Cand{0.0, 0x7fffffff, 0u};
            win = cand_warp_pick(win, lane < CS && win.i != 0x7fffffff);
        } else {
            win = lane < 8 ? s_wred[par][lane] : Cand{0.0, 0x7fffffff, 0u}; // every warp picks among the 8 partials itself
            win = cand_warp_pick(win, lane < 8 && win.i != 0x7fffffff);
        }
        int p = win.i;
        const REAL pivabs = fabs(win.v);
        if (p >= m) p = j; // nothing but NaNs: keep the diagonal, flagged singular below
        const REAL inv = (pivabs > 0.0 && pivabs < INFINITY) ? fast_rcp(win.v) : 0.0;
        if (cr == 0 && tid == 0) {
            ipiv[k0 + j] = k0 + p;
            s_piv[j] = p;
            if (!(pivabs > 0.0) && flags[FD_FLAG_SINGULAR] == 0) flags[FD_FLAG_SINGULAR] = k0 + j + 1;
            pmin = fmin(pmin, (double)pivabs);
            pmax = fmax(pmax, (double)pivabs);
        }
        // (3) the owners publish rows j and p in their own shared memory
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = cr * rows_per_cta + tid + q * 256;
            if (r == j) {
#pragma unroll
                for (int c = 0; c < NB; c += 2) *reinterpret_cast<REAL2*>(&s_pub[par][0][c]) = fd_make2(a[q][c], a[q][c + 1]);
            }
            if (r == p) {
#pragma unroll
                for (int c = 0; c < NB; c += 2) *reinterpret_cast<REAL2*>(&s_pub[par][1][c]) = fd_make2(a[q][c], a[q][c + 1]);
            }
        }
        const REAL* u;
        const REAL* oldj;
        if (CLUSTER) {
            cluster.sync();
            // (4) pull both rows from their owners
            if (tid < 2 * NB) {
                const int which = tid >> 5, c = tid & 31;
                const int owner = (which == 0 ? j : p) / rows_per_cta;
                s_row[par][which][c] = *cluster.map_shared_rank(&s_pub[par][which][c], owner);
            }
            __syncthreads();
            u = s_row[par][1]; // u[c] = U(j, j + c)
            oldj = s_row[par][0];
        } else {
            __syncthreads();
            u = s_pub[par][1];
            oldj = s_pub[par][0];
        }
        if (cr == 0 && tid < nb - j) G[(size_t)(j + tid) * lda + j] = u[tid]; // row j of U is final
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = cr * rows_per_cta + tid + q * 256;
            if (r == p && p != j) { // the row that sat at position j moves to position p
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    const REAL2 t = *reinterpret_cast<const REAL2*>(&oldj[c]);
                    a[q][c] = t.x;
                    a[q][c + 1] = t.y;
                }
            }
            if (r > j && r < m) {
                const REAL l = a[q][0] * inv;
                G[(size_t)j * lda + r] = l;
                REAL uu[NB];
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    const REAL2 t = *reinterpret_cast<const REAL2*>(&u[c]);
                    uu[c] = t.x;
                    uu[c + 1] = t.y;
                }
#pragma unroll
                for (int c = 1; c < NB; ++c) a[q][c - 1] = a[q][c] - l * uu[c]; // rank-1 update fused with the shift
                a[q][NB - 1] = 0.0;
            }
        }
    }
    if (cr == 0 && tid == 0) {
        pivstat[0] = fmin(pivstat[0], pmin);
        pivstat[1] = fmax(pivstat[1], pmax);
    }
    // every L entry is in global memory and visible; no CTA may exit while a peer can still read its shared memory
    if (CLUSTER) cluster.sync(); else __syncthreads();
    if (cr != 0) return;
    // Epilogue (CTA 0): the interchanges inside the panel's own finished columns.  Column c only sees the swaps that
    // came after it was finished (j > c).  The swaps touch at most 2 nb rows: gather that window with independent
    // loads, replay per column in shared memory, scatter back.
    if (warp == 0) {
        if (lane < nb) s_rows[lane] = lane;
        int cnt = nb;
        __syncwarp();
        for (int j = 0; j < nb; ++j) {
            const int p = s_piv[j];
            int ib;
            if (p < nb) {
                ib = p;
            } else {
                const bool hit = (nb + lane < cnt) && s_rows[nb + lane] == p;
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    ib = nb + __ffs(mask) - 1;
                } else {
                    if (lane == 0) s_rows[cnt] = p;
                    ib = cnt++;
                    __syncwarp();
                }
            }
            if (lane == 0) s_ib[j] = ib;
        }
        if (lane == 0) s_cnt = cnt;
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (warp == 1) {
        // the composition of all nb interchanges on the window, for k_lu_update: after the swaps, window entry e
        // holds what entry src[e] held before.  win = { cnt, rows[2 NB] (absolute), src[2 NB] }
        __shared__ int s_src[2 * NB];
        s_src[lane] = lane;
        s_src[lane + 32] = lane + 32;
        __syncwarp();
        if (lane == 0) {
            for (int j = 0; j < nb; ++j) {
                const int ib = s_ib[j];
                if (ib != j) {
                    const int t = s_src[j];
                    s_src[j] = s_src[ib];
                    s_src[ib] = t;
                }
            }
            win[0] = cnt;
        }
        __syncwarp();
        for (int e = lane; e < 2 * NB; e += 32) {
            win[1 + e] = e < cnt ? k0 + s_rows[e] : 0;
            win[1 + 2 * NB + e] = e < cnt ? s_src[e] : e;
        }
    }
    for (int t = tid; t < cnt * NB; t += 256) {
        const int e = t / NB, c = t % NB;
        s_fix[e][c] = c < nb ? G[(size_t)c * lda + s_rows[e]] : 0.0;
    }
    __syncthreads();
    if (tid < nb) {
        for (int j = tid + 1; j < nb; ++j) {
            const int ib = s_ib[j];
            if (ib != j) {
                const REAL t = s_fix[j][tid];
                s_fix[j][tid] = s_fix[ib][tid];
                s_fix[ib][tid] = t;
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < cnt * NB; t += 256) {
        const int e = t / NB, c = t % NB;
        // only the part of column c below the diagonal was stored by the column loop as L; rows <= c hold U
        if (c < nb && s_rows[e] > c) G[(size_t)c * lda + s_rows[e]] = s_fix[e][c];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_lu_update: everything that follows a panel, fused, one CTA per tile of 16 columns over ALL columns:
//   interchanges  (left of the panel: all nb swaps; the panel's own columns: the swaps that came after the column
//                  was finished; right of the panel: all nb swaps)
//   U12 = L11^-1 A12 and A22 -= L21 * U12 for the tiles right of the panel.
// The interchanges touch at most 2 nb distinct rows (the nb top rows and the pivot rows): warp 0 lists them once,
// the CTA gathers its 16 columns of those rows with independent loads, replays the swaps in shared memory, solves
// the unit-lower block with warp shuffles (lane = row) and scatters the rows back -- no chain of dependent global
// accesses.  The trailing update then streams L21 (coalesced, from L2) against the 32 x 16 U tile in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int UT = 16;            // columns per CTA
constexpr int UPD_THREADS = 256;

__global__ void __launch_bounds__(UPD_THREADS) k_lu_update(REAL* __restrict__ A, int lda, int n, int k0, int nb,
                                                           const int* __restrict__ win, int do_gemm, long long* dbg)
{
    __shared__ REAL s_vals[2 * NB][UT + 1]; // window rows x tile columns (after the interchanges)
    __shared__ __align__(16) REAL s_U[NB][UT]; // U12 tile, rows >= nb zero
    __shared__ REAL s_L[NB][NB + 1];        // L11 (strictly lower part)
    __shared__ int s_rows[2 * NB];
    __shared__ int s_src[2 * NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * UT;
    const bool is_left = c0 + UT <= k0;
    const bool is_right = c0 >= k0 + nb;
    const bool dbgt = dbg && tid == 0 && blockIdx.x == gridDim.x - 2;
    if (dbgt) dbg[0] = clock64();
    if (!is_left && !is_right) return; // the panel kernel finished its own columns
    const int cnt = win[0];
    if (tid < 2 * NB) {
        s_rows[tid] = win[1 + tid];
        s_src[tid] = win[1 + 2 * NB + tid];
    }
    if (is_right) {
        for (int t = tid; t < NB * NB; t += UPD_THREADS) {
            const int r = t % NB, c = t / NB;
            s_L[r][c] = (r < nb && c < nb && r > c) ? A[(size_t)(k0 + c) * lda + k0 + r] : 0.0;
        }
    }
    __syncthreads();
    if (dbgt) dbg[1] = clock64();
    // gather the window, already permuted: entry e receives the row that the composed interchanges bring there
    for (int t = tid; t < cnt * UT; t += UPD_THREADS) {
        const int e = t / UT, c = c0 + (t % UT);
        s_vals[e][t % UT] = c < n ? A[(size_t)c * lda + s_rows[s_src[e]]] : 0.0;
    }
    __syncthreads();
    if (dbgt) dbg[3] = dbg[2] = clock64();
    if (is_right) { // U12 tile = L11^-1 * top rows: lane = row, two columns per warp
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * warp + cc;
            REAL x = lane < nb ? s_vals[lane][c] : 0.0;
            for (int j = 0; j < nb; ++j) {
                const REAL xj = __shfl_sync(0xffffffffu, x, j);
                if (lane > j) x -= s_L[lane][j] * xj;
            }
            if (lane < nb) s_vals[lane][c] = x;
            s_U[lane][c] = lane < nb ? x : 0.0;
        }
        __syncthreads();
    }
    if (dbgt) dbg[4] = clock64();
    // scatter the window back
    for (int t = tid; t < cnt * UT; t += UPD_THREADS) {
        const int e = t / UT, c = c0 + (t % UT);
        if (c < n) A[(size_t)c * lda + s_rows[e]] = s_vals[e][t % UT];
    }
    if (!is_right || !do_gemm) return;
    __syncthreads(); // the pivot rows just written belong to the trailing matrix updated below
    if (dbgt) dbg[5] = clock64();
    // trailing update of this tile's columns: C[r][c] -= sum_k L21[r][k] * U12[k][c].  All 48 loads of a row are
    // issued before the first FMA (unconditional, clamped addresses; rows k >= nb of s_U are zero).
    const int r_begin = k0 + nb;
    for (int r = r_begin + tid; r < n; r += UPD_THREADS) {
        REAL l[NB], cv[UT];
#pragma unroll
        for (int k = 0; k < NB; ++k) l[k] = A[(size_t)(k0 + min(k, nb - 1)) * lda + r];
#pragma unroll
        for (int c = 0; c < UT; ++c) cv[c] = A[(size_t)min(c0 + c, n - 1) * lda + r];
        asm volatile("" ::: "memory");
#pragma unroll
        for (int k = 0; k < NB; ++k) {
#pragma unroll
            for (int c = 0; c < UT; c += 2) {
                const REAL2 u2 = *reinterpret_cast<const REAL2*>(&s_U[k][c]);
                cv[c] -= l[k] * u2.x;
                cv[c + 1] -= l[k] * u2.y;
            }
        }
#pragma unroll
        for (int c = 0; c < UT; ++c)
            if (c0 + c < n) A[(size_t)(c0 + c) * lda + r] = cv[c];
    }
    if (dbgt) dbg[6] = clock64();
}

// Trailing update for tall trailing matrices: C[m2 x m2] -= L21[m2 x 32] * U12[32 x m2] with both operands staged in
// shared memory (CTA tile 128 x 64, 8 x 4 outputs per thread), so L21 and U12 are read from L2 once per tile instead
// of once per 16 columns.  The C tile is loaded up front and written once.
constexpr int GM = 128, GN = 64;
__global__ void __launch_bounds__(256) k_lu_gemm(REAL* __restrict__ A, int lda, int n, int k0, int nb)
{
    __shared__ __align__(16) REAL s_a[NB][GM]; // L21 tile [k][row]
    __shared__ __align__(16) REAL s_b[NB][GN]; // U12 tile [k][col]
    const int r0 = k0 + nb + blockIdx.x * GM;
    const int c0 = k0 + nb + blockIdx.y * GN;
    const int tid = threadIdx.x;
    for (int t = tid; t < NB * GM; t += 256) {
        const int rr = t % GM, k = t / GM;
        s_a[k][rr] = (k < nb && r0 + rr < n) ? A[(size_t)(k0 + k) * lda + r0 + rr] : 0.0;
    }
    for (int t = tid; t < NB * GN; t += 256) {
        const int k = t % NB, cc = t / NB;
        s_b[k][cc] = (k < nb && c0 + cc < n) ? A[(size_t)(c0 + cc) * lda + k0 + k] : 0.0;
    }
    // thread tile: rows tr + 16 i (i < 8) so that the lanes of a warp touch consecutive rows (coalesced column-major
    // accesses, conflict-free shared-memory reads), columns tc .. tc + 3
    const int tr = tid % 16, tc = (tid / 16) * 4;
    REAL acc[8][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            acc[i][j] = A[(size_t)min(c0 + tc + j, n - 1) * lda + min(r0 + tr + 16 * i, n - 1)]; // C tile, clamped addresses
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < NB; ++k) {
        REAL a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = s_a[k][tr + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
            const REAL2 t = *reinterpret_cast<const REAL2*>(&s_b[k][tc + j]);
            b[j] = t.x;
            b[j + 1] = t.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] -= a[i] * b[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + tc + j;
        if (c >= n) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + tr + 16 * i;
            if (r < n) A[(size_t)c * lda + r] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Symmetric positive definite fast path (Gaussian kernel, uniform radius): LU WITHOUT pivoting.
// K + lambda I is SPD, so eliminating it in natural order is backward stable, and the Schur complement that the
// polynomial rows leave behind, -P^T K^-1 P, is negative definite -- no row interchange is ever needed for the
// saddle-point system [[K, P], [P^T, 0]].  Without a pivot search a block column needs no per-column reduction or
// barrier across the rows: every CTA factors the 32 x 32 diagonal block redundantly in shared memory, then each
// thread finishes one row of L21 (x U11^-1) or one column of U12 (L11^-1 x) on its own.  The factors have the same
// layout as the pivoted path (unit-lower L, U, identity permutation), so the solve kernels are shared.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NP_THREADS = 256;

__global__ void __launch_bounds__(NP_THREADS) k_lu_nopiv_panel(REAL* __restrict__ A, int lda, int n, int k0, int nb,
                                                               int* __restrict__ flags, double* __restrict__ pivstat)
{
    __shared__ __align__(16) REAL s_D[NB][NB + 2]; // diagonal block, row-major [r][c]; becomes L11 \ U11
    __shared__ REAL s_inv[NB];
    const int tid = threadIdx.x;
    for (int t = tid; t < NB * NB; t += NP_THREADS) {
        const int r = t % NB, c = t / NB;
        s_D[r][c] = (r < nb && c < nb) ? A[(size_t)(k0 + c) * lda + k0 + r] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    // LU of the diagonal block: step j updates the (r > j, c > j) entries with the un-scaled column j; the column
    // is scaled by 1/u_jj once at the end
    for (int j = 0; j < NB; ++j) {
        const REAL ujj = s_D[j][j];
        const REAL inv = (ujj != 0.0 && isfinite(ujj)) ? fast_rcp(ujj) : 0.0;
        if (tid == 0) s_inv[j] = inv;
        for (int t = tid; t < NB * NB; t += NP_THREADS) {
            const int r = t / NB, c = t % NB;
            if (r > j && c > j) s_D[r][c] -= (s_D[r][j] * inv) * s_D[j][c];
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0) {
        double pmin = pivstat[0], pmax = pivstat[1];
        for (int j = 0; j < nb; ++j) {
            const REAL v = fabs(s_D[j][j]);
            if (!(v > 0.0) || !isfinite(v)) {
                if (flags[FD_FLAG_SINGULAR] == 0) flags[FD_FLAG_SINGULAR] = k0 + j + 1;
            }
            pmin = fmin(pmin, (double)v);
            pmax = fmax(pmax, (double)v);
        }
        pivstat[0] = pmin;
        pivstat[1] = pmax;
    }
    for (int t = tid; t < NB * NB; t += NP_THREADS) {
        const int r = t / NB, c = t % NB;
        if (r > c) s_D[r][c] *= s_inv[c];
    }
    __syncthreads();
    const int m2 = n - k0 - nb; // rows below / columns right of the block
    const int row_ctas = (m2 + NP_THREADS - 1) / NP_THREADS;
    if (blockIdx.x == 0) { // write the factored diagonal block back
        for (int t = tid; t < NB * NB; t += NP_THREADS) {
            const int r = t % NB, c = t / NB;
            if (r < nb && c < nb) A[(size_t)(k0 + c) * lda + k0 + r] = s_D[r][c];
        }
    }
    if (m2 <= 0) return;
    if ((int)blockIdx.x < row_ctas) {
        // one row of L21 per thread: l = a U11^-1, i.e. forward substitution against the columns of U11
        const int r = k0 + nb + blockIdx.x * NP_THREADS + tid;
        if (r >= n) return;
        REAL x[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) x[c] = A[(size_t)(k0 + min(c, nb - 1)) * lda + r];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const REAL l = x[j] * s_inv[j];
            x[j] = l;
#pragma unroll
            for (int c = j + 1; c < NB; ++c) x[c] -= l * s_D[j][c];
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < nb) A[(size_t)(k0 + c) * lda + r] = x[c];
    } else {
        // one column of U12 per thread: u = L11^-1 a (unit lower)
        const int c = k0 + nb + (blockIdx.x - row_ctas) * NP_THREADS + tid;
        if (c >= n) return;
        REAL* col = A + (size_t)c * lda + k0;
        REAL x[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) x[r] = col[min(r, nb - 1)];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const REAL xj = x[j];
#pragma unroll
            for (int r = j + 1; r < NB; ++r) x[r] -= s_D[r][j] * xj;
        }
#pragma unroll
        for (int r = 0; r < NB; ++r)
            if (r < nb) col[r] = x[r];
    }
}

__global__ void k_lu_identity_perm(int n, int* __restrict__ ipiv, int* __restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        ipiv[i] = i;
        perm[i] = i;
    }
}

// perm[i] = original row that ends up in row i after all interchanges (single CTA, shared-memory resident)
__global__ void __launch_bounds__(256) k_lu_perm(const int* __restrict__ ipiv, int n, int* __restrict__ perm)
{
    extern __shared__ int s_perm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_perm[i] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < n; ++k) {
            const int p = ipiv[k];
            if (p != k) {
                const int t = s_perm[k];
                s_perm[k] = s_perm[p];
                s_perm[p] = t;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = s_perm[i];
}

__global__ void k_lu_init(int* flags, double* pivstat)
{
    if (threadIdx.x == 0) {
        flags[FD_FLAG_SINGULAR] = 0;
        flags[FD_FLAG_NONFINITE] = 0;
        pivstat[0] = INFINITY;
        pivstat[1] = 0.0;
    }
}

// (the namespace stays open: the launchers below are part of it)

template <int RPT, bool CLUSTER>
static cudaError_t launch_panel_cluster(cudaStream_t s, int cs, REAL* d_A, int lda, int n, int k0, int nb, int* d_ipiv,
                                        int* d_flags, double* d_pivstat, int* d_win)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_lu_panel_cluster<RPT, CLUSTER>, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win);
}

// window of a panel whose kernel did not export it (fallback panel): same list + composition, one warp
__global__ void __launch_bounds__(32) k_lu_window(const int* __restrict__ ipiv, int k0, int nb, int* __restrict__ win)
{
    __shared__ int s_rows[2 * NB], s_src[2 * NB], s_ib[NB];
    const int lane = threadIdx.x;
    if (lane < nb) s_rows[lane] = k0 + lane;
    const int my_piv = lane < nb ? ipiv[k0 + lane] : 0;
    int cnt = nb;
    __syncwarp();
    for (int j = 0; j < nb; ++j) {
        const int p = __shfl_sync(0xffffffffu, my_piv, j);
        int ib;
        if (p < k0 + nb) {
            ib = p - k0;
        } else {
            const bool hit = (nb + lane < cnt) && s_rows[nb + lane] == p;
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
                ib = nb + __ffs(mask) - 1;
            } else {
                if (lane == 0) s_rows[cnt] = p;
                ib = cnt++;
                __syncwarp();
            }
        }
        if (lane == 0) s_ib[j] = ib;
    }
    s_src[lane] = lane;
    s_src[lane + 32] = lane + 32;
    __syncwarp();
    if (lane == 0) {
        for (int j = 0; j < nb; ++j) {
            const int ib = s_ib[j];
            if (ib != j) {
                const int t = s_src[j];
                s_src[j] = s_src[ib];
                s_src[ib] = t;
            }
        }
        win[0] = cnt;
    }
    __syncwarp();
    for (int e = lane; e < 2 * NB; e += 32) {
        win[1 + e] = e < cnt ? s_rows[e] : 0;
        win[1 + 2 * NB + e] = e < cnt ? s_src[e] : e;
    }
}

cudaError_t launch_lu(fd_ctx* ctx, REAL* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                         double* d_pivstat, int* d_win)
{
    cudaStream_t s = ctx->stream;
    long long* d_dbg = ctx->d_lu_dbg;
    const bool want_dbg = d_dbg != nullptr;
    k_lu_init<<<1, 32, 0, s>>>(d_flags, d_pivstat);
    ctx->launches += 1;
    cudaError_t e = cudaSuccess;
    for (int k0 = 0; k0 < n && e == cudaSuccess; k0 += NB) {
        const int nb = min(NB, n - k0);
        const int m = n - k0;
        // smallest cluster (1, 2, 4, 8, 16 CTAs) x rows per thread (1..3) that holds the m panel rows in registers
        int cs = 0, rpt = 0;
        for (int c = 1; c <= 16 && !cs; c *= 2)
            for (int r = 1; r <= 3; ++r)
                if (m <= c * 256 * r) { cs = c; rpt = r; break; }
        if (cs) {
            if (cs == 1)
                e = rpt == 1 ? launch_panel_cluster<1, false>(s, 1, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win)
                  : rpt == 2 ? launch_panel_cluster<2, false>(s, 1, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win)
                             : launch_panel_cluster<3, false>(s, 1, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win);
            else
                e = rpt == 1 ? launch_panel_cluster<1, true>(s, cs, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win)
                  : rpt == 2 ? launch_panel_cluster<2, true>(s, cs, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win)
                             : launch_panel_cluster<3, true>(s, cs, d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat, d_win);
        } else { // taller than 12288 rows: panel through global memory, then its own interchanges are already applied
            k_lu_panel<false><<<1, PANEL_THREADS, 0, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
            k_lu_window<<<1, 32, 0, s>>>(d_ipiv, k0, nb, d_win);
            ctx->launches += 1;
            e = cudaGetLastError();
        }
        ctx->launches += 1;
        if (e != cudaSuccess) break;
        const int m2 = n - k0 - nb;
        const bool big = m2 >= 512; // tall trailing matrix: interchanges + TRSM fused, GEMM by the shared-memory-tiled kernel
        k_lu_update<<<(n + UT - 1) / UT, UPD_THREADS, 0, s>>>(d_A, lda, n, k0, nb, d_win, big ? 0 : 1, k0 == 0 ? d_dbg : nullptr);
        ctx->launches += 1;
        if (big) {
            dim3 grid((m2 + GM - 1) / GM, (m2 + GN - 1) / GN);
            k_lu_gemm<<<grid, 256, 0, s>>>(d_A, lda, n, k0, nb);
            ctx->launches += 1;
        }
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
    if (want_dbg) {
        long long h[8];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, d_dbg, 56, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[fd_lu] n=%d update(k0=0) cycles: list %lld gather %lld replay %lld trsm %lld scatter %lld gemm %lld\n", n,
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
    }
    if ((size_t)n * sizeof(int) > 64 * 1024) return cudaErrorInvalidValue;
    k_lu_perm<<<1, 256, (size_t)n * sizeof(int), s>>>(d_ipiv, n, d_perm);
    ctx->launches += 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Fused persistent no-pivot LU: the whole factorisation in ONE launch.
//
// For small systems the per-block-column launches above cost more than the arithmetic (N = 256: 17 launches of ~18 us
// for 12 Mflop).  Here a fixed set of CTAs walks all block columns and synchronises with a barrier instead of a launch
// boundary: one thread-block cluster (<= 16 CTAs, hardware barrier.cluster, ~0.3 us) for n <= 512, otherwise a
// cooperative grid of one CTA per SM with a release/acquire counter barrier in global memory.  The matrix stays in
// L2 / HBM (column-major, in place); all loads of it bypass L1 (ld.global.cg / cp.async.cg).
//
// Two-level blocking (outer block nbo, inner NB = 32) keeps the big trailing update compute bound: a rank-32 update
// of an m x m FP64 matrix moves 16 bytes per 64 flop (HBM bound above ~3000 rows); inside an outer block only the
// L-shaped border (column panel m x nbo, row panel nbo x m) is updated with K = 32, and the interior takes one
// C -= L21 * U12 with K = nbo.  nbo = 32 degenerates to the plain right-looking algorithm (empty L-shape).
//   per inner step : diagonal 32 x 32 block factored redundantly by warp 0 of every CTA (registers + shuffles,
//                    ~100 cycles per pivot: the critical path of the whole kernel), then one row of L21 (x U11^-1)
//                    or one column of U12 (L11^-1 x) per thread, barrier, L-shape update, barrier
//   per outer step : interior update (128 x 128 or 64 x 64 CTA tiles, 8 x 8 / 4 x 4 per thread, K chunks of 16
//                    double-buffered with cp.async), barrier
// The inverted diagonal blocks the slab solve wants (k_lu_invdiag) are produced on the way by two spare warps.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FZ_THREADS = 256;
constexpr int FZ_KC = 16;          // K chunk of the tile product
constexpr int FZ_TMAX = 128;       // largest CTA tile edge
#ifdef FD_LU_DMMA
// FP64: the tile product runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64).  Both stage strides are = 4 mod 16 doubles,
// which makes the fragment loads of a half warp (4 rows x 4 k) hit 16 different 8-byte banks.
constexpr int FZ_LDA = FZ_TMAX + 4; // row stride of the L21 stage ([k][row], rows contiguous)
constexpr int FZ_LDB = FZ_KC + 4;   // row stride of the U12 stage ([col][k], k contiguous)
#else
constexpr int FZ_LDA = FZ_TMAX;
constexpr int FZ_LDB = FZ_KC + 2;
#endif
constexpr int FZ_SMEM_REALS = 2 * FZ_KC * FZ_LDA + 2 * FZ_TMAX * FZ_LDB + 2 * NB * NB + NB;

__device__ __forceinline__ void fz_cp_async(REAL* smem_dst, const REAL* gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (sizeof(REAL2) == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fz_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void fz_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// barrier of the 256 threads of a fused-LU CTA (named barrier 1: barrier 0 separates the diagonal block from its solves)
__device__ __forceinline__ void fz_sync_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool CLUSTER>
__device__ __forceinline__ void fz_barrier(unsigned* counter, unsigned& target, unsigned G)
{
    if (CLUSTER) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        fz_sync_workers();
        if (threadIdx.x == 0) {
            target += G;
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while ((int)(v - target) < 0);
            __threadfence();
        }
        fz_sync_workers();
    }
}

#ifdef FD_LU_DMMA
// C[r0:r0+TM, c0:c0+TN] -= A[r0:, ka:kb] * A[ka:kb, c0:] on the FP64 tensor pipe; stores clipped to rows < r1, columns < c1.
// kb - ka is a multiple of 16, r0 / ka of 32.  Loads beyond the matrix are clamped to valid addresses (results dropped).
// 8 warps as 4 (rows) x 2 (columns); a warp owns (TM / 4) x (TN / 2) outputs = MT x NT accumulator tiles of 8 x 8 and
// loads MT + NT fragments per MT * NT DMMAs (TM = TN = 128: 12 loads per 32 DMMAs = 8192 FMAs -- the DFMA version read
// 16 bytes of shared memory per 16 FMAs and was bound by that).  The stage holds -L21 is not needed: the A fragment is
// negated in registers (4 DADD per 32 DMMAs).
template <int TM, int TN>
__device__ __forceinline__ void fz_gemm_tile(REAL* A, int lda, int n, int r0, int r1, int c0, int c1, int ka, int kb,
                                             REAL* s_a, REAL* s_b)
{
    constexpr int WR = TM / 4, WC = TN / 2, MT = WR / 8, NT = WC / 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, fr = lane >> 2, fk = lane & 3;
    double acc[MT][NT][2];
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
        const int c = c0 + wn * WC + ni * 8 + 2 * fk;
        const double* col0 = A + (size_t)min(c, n - 1) * lda;
        const double* col1 = A + (size_t)min(c + 1, n - 1) * lda;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const int r = min(r0 + wm * WR + mi * 8 + fr, lda - 1);
            acc[mi][ni][0] = __ldcg(col0 + r);
            acc[mi][ni][1] = __ldcg(col1 + r);
        }
    }
    auto issue = [&](int kc, int st) {
        REAL* sa = s_a + st * FZ_KC * FZ_LDA;
        for (int t = tid; t < FZ_KC * (TM / 2); t += FZ_THREADS) {
            const int kk = t / (TM / 2), q = t % (TM / 2);
            const int r = min(r0 + 2 * q, lda - 2);
            fz_cp_async(sa + kk * FZ_LDA + 2 * q, A + (size_t)(ka + kc + kk) * lda + r);
        }
        REAL* sb = s_b + st * FZ_TMAX * FZ_LDB;
        for (int t = tid; t < TN * (FZ_KC / 2); t += FZ_THREADS) {
            const int cc = t / (FZ_KC / 2), q = t % (FZ_KC / 2);
            const int c = min(c0 + cc, n - 1);
            fz_cp_async(sb + cc * FZ_LDB + 2 * q, A + (size_t)c * lda + ka + kc + 2 * q);
        }
        fz_cp_commit();
    };
    const int nch = (kb - ka) / FZ_KC;
    issue(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            issue((ch + 1) * FZ_KC, (ch + 1) & 1);
            fz_cp_wait<1>();
        } else {
            fz_cp_wait<0>();
        }
        fz_sync_workers();
        const double* sa = s_a + (ch & 1) * FZ_KC * FZ_LDA + fk * FZ_LDA + wm * WR + fr;
        const double* sb = s_b + (ch & 1) * FZ_TMAX * FZ_LDB + (wn * WC + fr) * FZ_LDB + fk;
#pragma unroll
        for (int k4 = 0; k4 < FZ_KC / 4; ++k4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) af[mi] = -sa[k4 * 4 * FZ_LDA + mi * 8];
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) bf[ni] = sb[ni * 8 * FZ_LDB + k4 * 4];
#pragma unroll
            for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                for (int ni = 0; ni < NT; ++ni)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                                 : "d"(af[mi]), "d"(bf[ni]));
        }
        fz_sync_workers();
    }
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
        const int c = c0 + wn * WC + ni * 8 + 2 * fk;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const int r = r0 + wm * WR + mi * 8 + fr;
            if (r < r1) {
                if (c < c1) __stcg(A + (size_t)c * lda + r, acc[mi][ni][0]);
                if (c + 1 < c1) __stcg(A + (size_t)(c + 1) * lda + r, acc[mi][ni][1]);
            }
        }
    }
}
#else
// C[r0:r0+TM, c0:c0+TN] -= A[r0:, ka:kb] * A[ka:kb, c0:]; stores clipped to rows < r1, columns < c1.  kb - ka is a
// multiple of 16, r0 / ka of 32.  Loads beyond the matrix are clamped to valid addresses (their results are dropped).
template <int TM, int TN>
__device__ __forceinline__ void fz_gemm_tile(REAL* A, int lda, int n, int r0, int r1, int c0, int c1, int ka, int kb,
                                             REAL* s_a, REAL* s_b)
{
    constexpr int TMt = TM / 16, TNt = TN / 16, MP = TMt / 2;
    const int tid = threadIdx.x, tr = tid & 15, tcg = tid >> 4;
    REAL acc[TMt][TNt];
#pragma unroll
    for (int j = 0; j < TNt; ++j) {
        const int c = min(c0 + tcg + 16 * j, n - 1);
#pragma unroll
        for (int i = 0; i < MP; ++i) {
            const int r = min(r0 + 2 * tr + 32 * i, lda - 2);
            const REAL2 v = __ldcg(reinterpret_cast<const REAL2*>(A + (size_t)c * lda + r));
            acc[2 * i][j] = v.x;
            acc[2 * i + 1][j] = v.y;
        }
    }
    auto issue = [&](int kc, int st) {
        REAL* sa = s_a + st * FZ_KC * FZ_LDA;
        for (int t = tid; t < FZ_KC * (TM / 2); t += FZ_THREADS) {
            const int kk = t / (TM / 2), q = t % (TM / 2);
            const int r = min(r0 + 2 * q, lda - 2);
            fz_cp_async(sa + kk * FZ_LDA + 2 * q, A + (size_t)(ka + kc + kk) * lda + r);
        }
        REAL* sb = s_b + st * FZ_TMAX * FZ_LDB;
        for (int t = tid; t < TN * (FZ_KC / 2); t += FZ_THREADS) {
            const int cc = t / (FZ_KC / 2), q = t % (FZ_KC / 2);
            const int c = min(c0 + cc, n - 1);
            fz_cp_async(sb + cc * FZ_LDB + 2 * q, A + (size_t)c * lda + ka + kc + 2 * q);
        }
        fz_cp_commit();
    };
    const int nch = (kb - ka) / FZ_KC;
    issue(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            issue((ch + 1) * FZ_KC, (ch + 1) & 1);
            fz_cp_wait<1>();
        } else {
            fz_cp_wait<0>();
        }
        fz_sync_workers();
        const REAL* sa = s_a + (ch & 1) * FZ_KC * FZ_LDA + 2 * tr;
        const REAL* sb = s_b + (ch & 1) * FZ_TMAX * FZ_LDB + tcg * FZ_LDB;
#pragma unroll 2
        for (int kk = 0; kk < FZ_KC; kk += 2) {
            REAL2 a0[MP], a1[MP], b[TNt];
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                a0[i] = *reinterpret_cast<const REAL2*>(sa + kk * FZ_LDA + 32 * i);
                a1[i] = *reinterpret_cast<const REAL2*>(sa + (kk + 1) * FZ_LDA + 32 * i);
            }
#pragma unroll
            for (int j = 0; j < TNt; ++j) b[j] = *reinterpret_cast<const REAL2*>(sb + 16 * j * FZ_LDB + kk);
#pragma unroll
            for (int i = 0; i < MP; ++i)
#pragma unroll
                for (int j = 0; j < TNt; ++j) {
                    acc[2 * i][j] -= a0[i].x * b[j].x;
                    acc[2 * i + 1][j] -= a0[i].y * b[j].x;
                }
#pragma unroll
            for (int i = 0; i < MP; ++i)
#pragma unroll
                for (int j = 0; j < TNt; ++j) {
                    acc[2 * i][j] -= a1[i].x * b[j].y;
                    acc[2 * i + 1][j] -= a1[i].y * b[j].y;
                }
        }
        fz_sync_workers();
    }
#pragma unroll
    for (int j = 0; j < TNt; ++j) {
        const int c = c0 + tcg + 16 * j;
        if (c >= c1) continue;
#pragma unroll
        for (int i = 0; i < MP; ++i) {
            const int r = r0 + 2 * tr + 32 * i;
            REAL* dst = A + (size_t)c * lda + r;
            if (r + 1 < r1)
                __stcg(reinterpret_cast<REAL2*>(dst), fd_make2(acc[2 * i][j], acc[2 * i + 1][j]));
            else if (r < r1)
                __stcg(dst, acc[2 * i][j]);
        }
    }
}

#endif

// One K range applied to up to two rectangular regions (the L-shaped border: column panel + row panel; or the
// interior alone), tiles dealt round-robin to the CTAs.  One call site per tile size keeps the kernel's code small:
// every phase runs once per block step, so code that does not fit the instruction cache runs at fetch speed.
struct FzRegions {
    int ra0, ra1, ca0, ca1; // region a: rows [ra0, ra1) x columns [ca0, ca1)
    int rb0, rb1, cb0, cb1; // region b
    int ka, kb;
    int sym;                // symmetric matrix: region a starts on the diagonal (ra0 == ca0) and only its tiles that touch
                            // the lower triangle (tile row >= tile column) are updated; region b is empty
};

// tiles of region a in the symmetric case: tile column j holds the tile rows j .. nrt - 1
__device__ __forceinline__ int fz_sym_tiles(int nrt, int nct)
{
    const int full = min(nrt, nct);
    return full * nrt - full * (full - 1) / 2;
}

template <int T>
__device__ __noinline__ void fz_update(REAL* A, int lda, int n, const FzRegions R, REAL* s_a, REAL* s_b)
{
    if (R.sym) {
        if (R.ra1 <= R.ra0 || R.ca1 <= R.ca0) return;
        const int nrt = (R.ra1 - R.ra0 + T - 1) / T, nct = (R.ca1 - R.ca0 + T - 1) / T;
        const int total = fz_sym_tiles(nrt, nct);
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            int j = 0, rem = t; // tile column j: nrt - j tiles
            while (rem >= nrt - j) { rem -= nrt - j; ++j; }
            const int i = j + rem;
            fz_gemm_tile<T, T>(A, lda, n, R.ra0 + i * T, R.ra1, R.ca0 + j * T, R.ca1, R.ka, R.kb, s_a, s_b);
        }
        return;
    }
    const int ta_r = R.ra1 > R.ra0 && R.ca1 > R.ca0 ? (R.ra1 - R.ra0 + T - 1) / T : 0;
    const int ta = ta_r * ((R.ca1 - R.ca0 + T - 1) / T);
    const int tb_r = R.rb1 > R.rb0 && R.cb1 > R.cb0 ? (R.rb1 - R.rb0 + T - 1) / T : 0;
    const int tb = tb_r * ((R.cb1 - R.cb0 + T - 1) / T);
    for (int t = blockIdx.x; t < ta + tb; t += gridDim.x) {
        int r0, r1, c0, c1;
        if (t < ta) {
            r0 = R.ra0 + (t % ta_r) * T, r1 = R.ra1, c0 = R.ca0 + (t / ta_r) * T, c1 = R.ca1;
        } else {
            r0 = R.rb0 + ((t - ta) % tb_r) * T, r1 = R.rb1, c0 = R.cb0 + ((t - ta) / tb_r) * T, c1 = R.cb1;
        }
        fz_gemm_tile<T, T>(A, lda, n, r0, r1, c0, c1, R.ka, R.kb, s_a, s_b);
    }
}

__device__ __forceinline__ int fz_tiles(const FzRegions& R, int T)
{
    if (R.sym) {
        if (R.ra1 <= R.ra0 || R.ca1 <= R.ca0) return 0;
        return fz_sym_tiles((R.ra1 - R.ra0 + T - 1) / T, (R.ca1 - R.ca0 + T - 1) / T);
    }
    const int ta = R.ra1 > R.ra0 && R.ca1 > R.ca0 ? ((R.ra1 - R.ra0 + T - 1) / T) * ((R.ca1 - R.ca0 + T - 1) / T) : 0;
    const int tb = R.rb1 > R.rb0 && R.cb1 > R.cb0 ? ((R.rb1 - R.rb0 + T - 1) / T) * ((R.cb1 - R.cb0 + T - 1) / T) : 0;
    return ta + tb;
}

// x <- x * U11^-1 for a row vector held by one thread (s_U[j][c] = U11[j][c], s_inv[j] = 1 / u_jj); fully unrolled,
// static register indices, the row of U11 is a broadcast read.
__device__ __forceinline__ void fz_row_solve(REAL (&x)[NB], const REAL* s_U, const REAL* s_inv)
{
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const REAL l = x[j] * s_inv[j];
        x[j] = l;
#pragma unroll
        for (int c = 0; c < NB; c += 2) {
            if (c + 1 > j) { // pair (c, c+1): at least c+1 lies right of the pivot column
                const REAL2 u = *reinterpret_cast<const REAL2*>(s_U + j * NB + c);
                if (c > j) x[c] -= l * u.x;
                x[c + 1] -= l * u.y;
            }
        }
    }
}

// x <- L11^-1 x for a column vector held by one thread (s_Lt[j][r] = L11[r][j], unit diagonal)
__device__ __forceinline__ void fz_col_solve(REAL (&x)[NB], const REAL* s_Lt)
{
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const REAL xj = x[j];
#pragma unroll
        for (int r = 0; r < NB; r += 2) {
            if (r + 1 > j) {
                const REAL2 l2 = *reinterpret_cast<const REAL2*>(s_Lt + j * NB + r);
                if (r > j) x[r] -= l2.x * xj;
                x[r + 1] -= l2.y * xj;
            }
        }
    }
}

__device__ __forceinline__ REAL fz_safe_rcp(REAL p) { return (p != (REAL)0 && isfinite(p)) ? fast_rcp(p) : (REAL)0; }

template <bool CLUSTER>
__global__ void __launch_bounds__(FZ_THREADS, 1) k_lu_nopiv_fused(REAL* A, int lda, int n, int nbo, int* ipiv, int* perm,
                                                                  int* flags, double* pivstat, double* Tinv,
                                                                  unsigned* sync_counter, unsigned sync_base, int dbg, int sym)
{
    extern __shared__ __align__(16) unsigned char fz_smem_raw[];
    long long tprobe[12];
    int tstep = 0;
#define FZ_PROBE(i) do { if (dbg && blockIdx.x == 0 && threadIdx.x == 0 && tstep == dbg) tprobe[i] = clock64(); } while (0)
    REAL* s_a = reinterpret_cast<REAL*>(fz_smem_raw);
    REAL* s_b = s_a + 2 * FZ_KC * FZ_LDA;
    REAL* s_U = s_b + 2 * FZ_TMAX * FZ_LDB; // [NB][NB]: U11 (row j valid from column j & ~1 on)
    REAL* s_Lt = s_U + NB * NB;             // [NB][NB]: s_Lt[j][r] = L11[r][j]
    REAL* s_inv = s_Lt + NB * NB;           // 1 / u_jj
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned G = gridDim.x;
    unsigned target = sync_base;
    double pmin = INFINITY, pmax = 0.0;
    int singular = 0;
    if (blockIdx.x == 0)
        for (int i = tid; i < n; i += FZ_THREADS) {
            ipiv[i] = i;
            perm[i] = i;
        }
    for (int K0 = 0; K0 < n; K0 += nbo) {
        const int Kend = min(K0 + nbo, n);
#pragma unroll 1
        for (int k0 = K0; k0 < Kend; k0 += NB) {
            const int nb = min(NB, n - k0);
            const int ks = k0 + nb, m2 = n - ks;
            const int ngr = (m2 + 31) >> 5; // groups of 32 rows of L21; as many groups of 32 columns of U12
            ++tstep;
            FZ_PROBE(0);
            // The last block has no barrier behind it: CTA 0 alone factors it, inverts it and writes it back (a second CTA
            // reading the raw block could find CTA 0's factors there).
            if (m2 <= 0 && blockIdx.x != 0) break;
            // Warp 0 factors the diagonal block and takes no L21 / U12 group; the branches are exclusive so that the
            // preloaded x[] is not live (no registers reserved) inside the latency-critical pivot sequence.
            if (warp == 0) {
                FZ_PROBE(1);
                // ---- diagonal block: lane = row, a[c] = A[lane][c]; fully unrolled (static register indices).  Row j
                // is final at step j: its lane publishes it in shared memory, the rows below read it back as a
                // broadcast.  The next pivot's reciprocal is started as soon as its element is updated, so the
                // MUFU + Newton chain overlaps the rest of the rank-1 update.
                REAL a[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    a[c] = (lane < nb && c < nb) ? __ldcg(A + (size_t)(k0 + c) * lda + k0 + lane) : (lane == c ? (REAL)1 : (REAL)0);
                REAL mypiv = 1;
                REAL p = __shfl_sync(0xffffffffu, a[0], 0);
                REAL inv = fz_safe_rcp(p);
                FZ_PROBE(9);
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if (lane == j) {
                        mypiv = p;
                        s_inv[j] = inv;
#pragma unroll
                        for (int c = 0; c < NB; c += 2)
                            if (c + 1 >= j) *reinterpret_cast<REAL2*>(s_U + j * NB + c) = fd_make2(a[c], a[c + 1]);
                    }
                    const REAL l = a[j] * inv; // meaningful for lanes > j; finished rows carry don't-care values
                    s_Lt[j * NB + lane] = lane > j ? l : (lane == j ? (REAL)1 : (REAL)0);
                    __syncwarp();
                    if (j + 1 < NB) {
                        a[j + 1] -= l * s_U[j * NB + j + 1];
                        p = __shfl_sync(0xffffffffu, a[j + 1], j + 1);
                        inv = fz_safe_rcp(p);
#pragma unroll
                        for (int c = 0; c < NB; c += 2) {
                            if (c + 1 > j + 1) {
                                const REAL2 u = *reinterpret_cast<const REAL2*>(s_U + j * NB + c);
                                if (c > j + 1) a[c] -= l * u.x;
                                a[c + 1] -= l * u.y;
                            }
                        }
                    }
                }
                __syncwarp();
                FZ_PROBE(10);
                if (blockIdx.x == 0) {
                    const double v = lane < nb ? fabs((double)mypiv) : 1.0;
                    const bool bad = lane < nb && (!(v > 0.0) || !isfinite(v));
                    const unsigned badmask = __ballot_sync(0xffffffffu, bad);
                    if (badmask && !singular) singular = k0 + __ffs(badmask);
                    double lo = lane < nb ? v : INFINITY, hi = lane < nb ? v : 0.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                    }
                    pmin = fmin(pmin, lo);
                    pmax = fmax(pmax, hi);
                }
                FZ_PROBE(2);
                asm volatile("bar.sync 0;" ::: "memory");
                FZ_PROBE(3);
            } else {
                // ---- L21 = A21 U11^-1 (one row per thread), U12 = L11^-1 A12 (one column per thread), in groups of
                // 32 dealt round-robin to the warps 1..7 of all CTAs; the first group's loads are in flight while
                // warp 0 factors.  Two more "groups" apply the same two solves to the identity: the inverses of U11
                // and L11 that the slab solve (fd_solve.cu) multiplies with.
                // A symmetric matrix (the no-pivot LU is only taken for symmetric systems) has U = D L^T: U12 is the scaled
                // transpose of L21, so the column groups disappear -- every row group writes its 32 rows of L21 and, scaled
                // by the pivots, the matching 32 columns of U12.  Only the lower triangle of the trailing matrix is
                // ever updated then (the tile sets below), half the flops of the factorisation.
                REAL x[NB];
                const int ncg = sym ? 0 : ngr;          // column groups
                const int g_inv = ngr + ncg;            // first of the two inverse "groups"
                const int ngroups = g_inv + (Tinv != nullptr ? 2 : 0);
                int g = m2 <= 0 ? warp - 1 : (int)blockIdx.x + (int)G * (warp - 1);
                auto load_group = [&](int gg) {
                    if (gg < ngr) {
                        const int r = min(ks + 32 * gg + lane, n - 1);
#pragma unroll
                        for (int c = 0; c < NB; ++c) x[c] = __ldcg(A + (size_t)(k0 + c) * lda + r);
                    } else if (gg < g_inv) {
                        const int c = min(ks + 32 * (gg - ngr) + lane, n - 1);
                        const REAL2* src = reinterpret_cast<const REAL2*>(A + (size_t)c * lda + k0);
#pragma unroll
                        for (int r = 0; r < NB; r += 2) {
                            const REAL2 v = __ldcg(src + r / 2);
                            x[r] = v.x;
                            x[r + 1] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < NB; ++c) x[c] = c == lane ? (REAL)1 : (REAL)0;
                    }
                };
                if (g < ngroups) load_group(g);
                asm volatile("bar.sync 0;" ::: "memory");
#pragma unroll 1
                for (bool first = true; g < ngroups; g += (int)G * (FZ_THREADS / 32 - 1), first = false) {
                    if (!first) load_group(g);
                    const bool row_kind = g < ngr || g == g_inv;
                    if (row_kind)
                        fz_row_solve(x, s_U, s_inv);
                    else
                        fz_col_solve(x, s_Lt);
                    if (g < ngr) {
                        const int r = ks + 32 * g + lane;
                        if (r < n) {
#pragma unroll
                            for (int c = 0; c < NB; ++c) __stcg(A + (size_t)(k0 + c) * lda + r, x[c]);
                            if (sym) { // U12[c][r] = u_cc * L21[r][c]: column r of the row panel, 32 contiguous values
                                REAL2* dst = reinterpret_cast<REAL2*>(A + (size_t)r * lda + k0);
#pragma unroll
                                for (int c = 0; c < NB; c += 2)
                                    __stcg(dst + c / 2, fd_make2(x[c] * s_U[c * NB + c], x[c + 1] * s_U[(c + 1) * NB + c + 1]));
                            }
                        }
                    } else if (g < g_inv) {
                        const int c = ks + 32 * (g - ngr) + lane;
                        if (c < n) {
                            REAL2* dst = reinterpret_cast<REAL2*>(A + (size_t)c * lda + k0);
#pragma unroll
                            for (int r = 0; r < NB; r += 2) __stcg(dst + r / 2, fd_make2(x[r], x[r + 1]));
                        }
                    } else if (g == g_inv) { // x[c] = U11^-1[lane][c]; stored transposed: out[c * 32 + r] = inverse[r][c]
                        double* out = Tinv + ((size_t)(k0 / NB) * 2 + 1) * NB * NB;
#pragma unroll
                        for (int c = 0; c < NB; ++c) out[c * NB + lane] = (double)x[c];
                    } else { // x[r] = L11^-1[r][lane]
                        double* out = Tinv + ((size_t)(k0 / NB) * 2) * NB * NB + lane * NB;
#pragma unroll
                        for (int r = 0; r < NB; ++r) out[r] = (double)x[r];
                    }
                }
            }
            // The factored diagonal block goes back to the matrix only AFTER the step's grid barrier: every CTA reads the raw
            // block from global memory at the top of the step (warp 0), and a CTA that leaves the previous barrier late must
            // not find CTA 0's factors there (seen as run-to-run differences of L21 at n >= 4096).  Nothing inside the kernel
            // reads the block again; s_U / s_Lt stay valid until warp 0 starts the next step, one barrier further on.
            auto write_back_diag = [&]() {
                if (blockIdx.x == 0 && warp == 5 && lane < nb) {
#pragma unroll 1
                    for (int c = 0; c < nb; ++c)
                        __stcg(A + (size_t)(k0 + c) * lda + k0 + lane, c >= lane ? s_U[lane * NB + c] : s_Lt[c * NB + lane]);
                }
            };
            if (m2 <= 0) {
                write_back_diag(); // last block: CTA 0 is the only reader
                break;
            }
            FZ_PROBE(5);
            fz_barrier<CLUSTER>(sync_counter, target, G);
            FZ_PROBE(6);
            write_back_diag();
            // ---- L-shaped border of the outer block, K = 32
            if (ks < Kend) {
                const FzRegions R = {ks, n, ks, Kend, ks, Kend, Kend, n, k0, ks, sym};
                if (fz_tiles(R, 128) >= 2 * (int)G)
                    fz_update<128>(A, lda, n, R, s_a, s_b);
                else
                    fz_update<64>(A, lda, n, R, s_a, s_b);
                fz_barrier<CLUSTER>(sync_counter, target, G);
            }
        }
        if (Kend < n) {
            // ---- interior of the trailing matrix, K = nbo
            const FzRegions R = {Kend, n, Kend, n, 0, 0, 0, 0, K0, Kend, sym};
            if (fz_tiles(R, 128) >= (int)G)
                fz_update<128>(A, lda, n, R, s_a, s_b);
            else
                fz_update<64>(A, lda, n, R, s_a, s_b);
            FZ_PROBE(7);
            fz_barrier<CLUSTER>(sync_counter, target, G);
            FZ_PROBE(8);
        }
    }
    if (dbg && blockIdx.x == 0 && tid == 0 && tstep >= dbg)
        printf("[fz] n=%d step %d cycles: diag %lld (loads %lld loop %lld tail %lld) sync %lld bar1 %lld update %lld bar2 %lld\n", n, dbg,
               tprobe[2] - tprobe[1], tprobe[9] - tprobe[1], tprobe[10] - tprobe[9], tprobe[2] - tprobe[10], tprobe[3] - tprobe[2],
               tprobe[6] - tprobe[5], tprobe[7] - tprobe[6], tprobe[8] - tprobe[7]);
#undef FZ_PROBE
    if (blockIdx.x == 0 && tid == 0) {
        flags[FD_FLAG_SINGULAR] = singular;
        flags[FD_FLAG_NONFINITE] = 0;
        pivstat[0] = pmin;
        pivstat[1] = pmax;
    }
}

// (A look-ahead variant -- a ninth warp per CTA factoring the next diagonal block while the workers run the trailing update --
// was built and measured in round 2 and LOST on B200: n = 260 0.205 ms against 0.152 ms.  The diagonal warp's 32-pivot
// dependent chain shares its SM sub-partition's FP64 pipe with the workers' DMMAs, which hold the pipe 16 cycles each; the
// chain stretched from ~300 to ~940 cycles per pivot and stayed the critical path.  Probes: profiles/r2d_lu_probe.log;
// the code is in the history at commit 648e060.)

// number of barriers k_lu_nopiv_fused executes for (n, nbo): the host advances the counter base by G times this
static unsigned fz_barrier_count(int n, int nbo)
{
    unsigned cnt = 0;
    for (int K0 = 0; K0 < n; K0 += nbo) {
        const int Kend = min(K0 + nbo, n);
        for (int k0 = K0; k0 < Kend; k0 += NB) {
            const int nb = min(NB, n - k0), ks = k0 + nb;
            if (n - ks <= 0) break;
            cnt += 1;
            if (ks < Kend) cnt += 1;
        }
        if (Kend < n) cnt += 1;
    }
    return cnt;
}

cudaError_t launch_lu_nopivot_fused(fd_ctx* ctx, REAL* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                    double* d_pivstat, double* d_Tinv, int sym)
{
    if (ctx->dbg.lu_sym_off) sym = 0;
    const size_t smem = (size_t)FZ_SMEM_REALS * sizeof(REAL);
    const fd_debug_opts& o = ctx->dbg;
    int dbg = o.lu_debug;
    const int cluster_max_n = o.lu_cluster_max_n ? o.lu_cluster_max_n : 512;
    int nbo = o.lu_nbo ? o.lu_nbo : (n <= 1536 ? 32 : (n <= 3072 ? 128 : 256));
    nbo = max(NB, nbo / NB * NB);
    cudaStream_t s = ctx->stream;
    unsigned base = ctx->sync_base;
    if (n <= cluster_max_n) {
        // one cluster: 4 CTAs up to n = 64, 8 up to 128, else 16
        const int cs = o.lu_cluster ? o.lu_cluster : (n <= 64 ? 4 : (n <= 128 ? 8 : 16));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(FZ_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ctx->launches += 1;
        return cudaLaunchKernelEx(&cfg, k_lu_nopiv_fused<true>, d_A, lda, n, nbo, d_ipiv, d_perm, d_flags, d_pivstat, d_Tinv,
                                  ctx->d_sync, base, dbg, sym);
    }
    const int G = ctx->sm_count;
    void* args[] = {&d_A, &lda, &n, &nbo, &d_ipiv, &d_perm, &d_flags, &d_pivstat, &d_Tinv, &ctx->d_sync, &base, &dbg, &sym};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_lu_nopiv_fused<false>, dim3(G), dim3(FZ_THREADS), args, smem, s);
    if (e == cudaSuccess) { // the counter only moves when the kernel that advances it was really enqueued
        ctx->sync_base = base + (unsigned)G * fz_barrier_count(n, nbo);
        ctx->launches += 1;
    }
    return e;
}

// per-device function attributes of this instantiation's kernels (called from fd_ctx_create through fd_factor_setup)
cudaError_t setup_attributes()
{
    cudaError_t e = cudaFuncSetAttribute(k_lu_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_perm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_panel_cluster<1, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_panel_cluster<2, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_panel_cluster<3, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int smem = (int)((size_t)FZ_SMEM_REALS * sizeof(REAL));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_nopiv_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_nopiv_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lu_nopiv_fused<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

// LU without pivoting for the symmetric positive definite case (see k_lu_nopiv_panel)
cudaError_t launch_lu_nopivot(fd_ctx* ctx, REAL* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                 double* d_pivstat)
{
    cudaStream_t s = ctx->stream;
    k_lu_init<<<1, 32, 0, s>>>(d_flags, d_pivstat);
    k_lu_identity_perm<<<(n + 255) / 256, 256, 0, s>>>(n, d_ipiv, d_perm);
    ctx->launches += 2;
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);
        const int m2 = n - k0 - nb;
        const int ctas = m2 > 0 ? 2 * ((m2 + NP_THREADS - 1) / NP_THREADS) : 1;
        k_lu_nopiv_panel<<<ctas, NP_THREADS, 0, s>>>(d_A, lda, n, k0, nb, d_flags, d_pivstat);
        ctx->launches += 1;
        if (m2 > 0) {
            dim3 grid((m2 + GM - 1) / GM, (m2 + GN - 1) / GN);
            k_lu_gemm<<<grid, 256, 0, s>>>(d_A, lda, n, k0, nb);
            ctx->launches += 1;
        }
    }
    return cudaGetLastError();
}

} // namespace FD_LU_NS
