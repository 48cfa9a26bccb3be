// mcsce_b200 - A7: hydrophobic potential of mean force on the Hasel-Hendrickson-Still surface, sm_100a.
//
//   sasa_kernel        SA_i = S_i * prod_{j != i, d_ij < R_i + R_j + 2 R_s} (1 - P_i P_ij b_ij / S_i)      (libsasa.py:89-108)
//   hpmf_pair_kernel   V = 1/2 sum_{i != j carbons, SA > A_c} tanh(SA_i) tanh(SA_j) sum_k H_k exp(-((r - C_k)/W_k)^2)
//                                                                                                          (libhpmf.py:92-109)
//   hpmf_finish_kernel fixed-order sum of the per-block partial sums
//   hpmf_trial_kernel  V(environment + one trial side chain) for every (trial, conformer) of a growth step from the
//                      environment's areas, weights and pair fields - the environment is evaluated once, not per trial
//
// One structure per blockIdx.y, 128 atoms (carbons) per block, partner atoms staged through shared memory in tiles
// of 128.  Numerics follow the reference: pair difference, squared norm and square root in float32 (np.linalg.norm
// of a float32 array: x*x, add.reduce over three elements, sqrt), everything after that in float64, the product
// over partners in ascending partner order.  Pairs that cannot pass the distance test are rejected on the float32
// squared distance; Gaussian terms beyond HPMF_CUTOFF (14 A: < 4.5e-10 each) are dropped.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/mcsce_b200.h"

int mcsce_fail_shared(const std::string &m);
#define HCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return mcsce_fail_shared(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

#define HT 128                 // threads per block = atoms per tile
#define MAXB MCSCE_SASA_MAX_BONDS
#define MAX_SASA_TYPES 32
#ifndef HPMF_CUTOFF
#define HPMF_CUTOFF 14.0f      // carbon pairs beyond this are dropped: each term is below 0.091 exp(-((14 - 7.117)/1.574)^2) = 4.5e-10,
                               // a few thousand of them stay five orders of magnitude under the 1e-3 tolerance (18 A before:
                               // < 2e-22 each, twice the pairs - the Gaussian sums were 40 % of the per-trial kernel, ncu)
#endif
#define ABSENT_X 1.0e14f       // x beyond this = atom not placed (the growth parks unplaced atoms at 1e15)

struct HpmfDev {
    int n_atoms, n_typed, n_types, n_carbon;
    const int *typed_atom, *type_of, *bonded, *carbon;   // carbon: [n_carbon] typed index
    const double *thr;                                   // [n_types*n_types] R_i + R_j + 2 R_s
    const double *R, *P, *piRRs, *S;                     // per type: R, P, pi (R + R_s), 4 pi (R + R_s)^2
    double pij_bonded, pij_non_bonded;
    float pre2;                                          // float32 prefilter on d^2: (largest threshold)^2, padded
    double C[3], W[3], H[3], sa_threshold;
    float Cf[3], iWf[3], Hf[3];                          // float32 copies for the Gaussians (1/W)
};

__device__ __forceinline__ float norm2_f32(float ax, float ay, float az, float bx, float by, float bz) {
    float x = __fsub_rn(ax, bx), y = __fsub_rn(ay, by), z = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// fused form for range tests only: never feeds a result
__device__ __forceinline__ float dist2_fast(float ax, float ay, float az, float bx, float by, float bz) {
    const float x = ax - bx, y = ay - by, z = az - bz;
    return fmaf(z, z, fmaf(y, y, x * x));
}

// one factor of the product: atom i (P_i, S_i, R_i) shadowed by partner j (pi (R_j + R_s), R_j) at distance d < thr:
//   1 - P_i p_ij b_ij / S_i ,  b_ij = pi (R_j + R_s) (thr - d) (1 + (R_j - R_i) / (d + 1e-8))          (libsasa.py:100-105)
// The two float64 divisions (by d and by S_i) were 2/3 of the instructions of the per-trial kernels; the reciprocal of d comes
// from the float32 MUFU result refined by two Newton steps in float64 (relative error < 1e-15), 1/S_i is a per-type constant
// (one more rounding, 1e-16).  Surface areas stay within 1e-12 relative of the reference's (the tests ask 1e-9).
__device__ __forceinline__ double recip_f64(double x) {
    double y = (double)__frcp_rn((float)x);
    y = y * (2.0 - x * y);
    return y * (2.0 - x * y);
}
__device__ __forceinline__ double sasa_term(double Pi, double Si, double Ri, double piRRs_j, double Rj, double thr, float d,
                                            double pij) {
    // dij + 1e-8 stays float32 in the reference (float32 array + Python scalar)
    const double dd = (double)__fadd_rn(d, 1e-8f);
    const double bij = (piRRs_j * (thr - (double)d)) * (1.0 + (Rj - Ri) * recip_f64(dd));
    return 1.0 - ((Pi * pij) * bij) * recip_f64(Si);
}
// both factors of a pair at once: (i shadowed by j, j shadowed by i); invS = 1 / S per type
__device__ __forceinline__ void sasa_term_pair(double Pi, double invSi, double Ri, double piRRs_i, double Pj, double invSj, double Rj,
                                               double piRRs_j, double thr, float d, double pij, double &t_ij, double &t_ji) {
    const double dd = (double)__fadd_rn(d, 1e-8f);
    const double gap = thr - (double)d, q = (Rj - Ri) * recip_f64(dd);
    t_ij = 1.0 - ((Pi * pij) * ((piRRs_j * gap) * (1.0 + q))) * invSi;
    t_ji = 1.0 - ((Pj * pij) * ((piRRs_i * gap) * (1.0 - q))) * invSj;
}
// tanh of a surface area: exactly 1 in float64 beyond 19.07 (1 - tanh x < 2^-54), which most exposed carbons are
__device__ __forceinline__ double tanh_area(double a) { return a > 19.07 ? 1.0 : tanh(a); }

// nt_lim: only the first nt_lim typed atoms exist (a partially grown structure); n_typed for a whole one
__global__ void __launch_bounds__(HT) sasa_kernel(HpmfDev s, const float *__restrict__ coords, double *__restrict__ sasa,
                                                  int nt_lim) {
    __shared__ float4 sp[HT];               // x, y, z, type (int bits): one 16-byte load per partner
    __shared__ double s_thr[MAX_SASA_TYPES * MAX_SASA_TYPES];
    __shared__ double s_R[MAX_SASA_TYPES], s_piRRs[MAX_SASA_TYPES];
    const int tid = threadIdx.x, nt = s.n_types;
    const long long base = (long long)blockIdx.y * s.n_atoms;
    if (blockIdx.x * HT >= nt_lim) {            // whole block beyond the atoms that exist
        const int i0 = blockIdx.x * HT + tid;
        if (i0 < s.n_typed) sasa[(long long)blockIdx.y * s.n_typed + i0] = 0.0;
        return;
    }
    for (int k = tid; k < nt * nt; k += HT) s_thr[k] = s.thr[k];
    if (tid < nt) { s_R[tid] = s.R[tid]; s_piRRs[tid] = s.piRRs[tid]; }
    const int i = blockIdx.x * HT + tid;
    const bool live = i < nt_lim;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    int ti = 0, b0 = -1, b1 = -1, b2 = -1, b3 = -1;
    double Ri = 0., Pi = 0., Si = 1.;
    if (live) {
        const float *p = coords + (base + s.typed_atom[i]) * 3;
        xi = p[0]; yi = p[1]; zi = p[2];
        ti = s.type_of[i];
        b0 = s.bonded[i * MAXB + 0]; b1 = s.bonded[i * MAXB + 1];
        b2 = s.bonded[i * MAXB + 2]; b3 = s.bonded[i * MAXB + 3];
        Ri = s.R[ti]; Pi = s.P[ti]; Si = s.S[ti];
    }
    double prod = 1.0;
    for (int j0 = 0; j0 < nt_lim; j0 += HT) {
        __syncthreads();
        const int jl = j0 + tid;
        if (jl < nt_lim) {
            const float *p = coords + (base + s.typed_atom[jl]) * 3;
            sp[tid] = make_float4(p[0], p[1], p[2], __int_as_float(s.type_of[jl]));
        }
        __syncthreads();
        if (!live) continue;
        const int n = min(HT, nt_lim - j0);
        const double *thr_row = s_thr + ti * nt;
#pragma unroll 4
        for (int jj = 0; jj < n; ++jj) {
            const float4 q = sp[jj];
            if (dist2_fast(xi, yi, zi, q.x, q.y, q.z) < s.pre2) {     // superset test (pre2 is padded by 1e-3)
                const int j = j0 + jj;
                if (j == i) continue;
                const float d = __fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z));
                const int tj = __float_as_int(q.w);
                const double thr = thr_row[tj];
                if ((double)d < thr) {
                    const double pij = (j == b0 || j == b1 || j == b2 || j == b3) ? s.pij_bonded : s.pij_non_bonded;
                    prod = __dmul_rn(prod, sasa_term(Pi, Si, Ri, s_piRRs[tj], s_R[tj], thr, d, pij));
                }
            }
        }
    }
    // an atom parked at 1e15 (not placed yet, the growth's convention) has no surface and no partners
    if (live) sasa[(long long)blockIdx.y * s.n_typed + i] = xi > ABSENT_X ? 0.0 : __dmul_rn(Si, prod);
    else if (i < s.n_typed) sasa[(long long)blockIdx.y * s.n_typed + i] = 0.0;
}

// Sum of the three Gaussians of one carbon pair.  The reference evaluates them in float32 (rij is a float32 array and the
// Python scalars do not promote it, libhpmf.py:101-104), so float32 it is here too: (r - C)/W, square, exp and the weighted
// sum in float32 (MUFU.EX2 through __expf), the result promoted to float64 for the weight products and the sums.  Against the
// float64 form this kernel used before: 1e-7 relative on V, one tenth of the instructions.
__device__ __forceinline__ double gauss3(const HpmfDev &s, float r) {
    const float t0 = (r - s.Cf[0]) * s.iWf[0], t1 = (r - s.Cf[1]) * s.iWf[1], t2 = (r - s.Cf[2]) * s.iWf[2];
    return (double)(s.Hf[0] * __expf(-(t0 * t0)) + s.Hf[1] * __expf(-(t1 * t1)) + s.Hf[2] * __expf(-(t2 * t2)));
}

// nc_lim: only the first nc_lim carbons exist.  w_out / f_out (optional, [n_struct*n_carbon]): the weight of every
// carbon and its field F_i = sum_{j != i} w_j G_ij - what the per-trial kernel builds its increments on.
__global__ void __launch_bounds__(HT) hpmf_pair_kernel(HpmfDev s, const float *__restrict__ coords,
                                                       const double *__restrict__ sasa, double *__restrict__ partial,
                                                       int nc_lim, double *__restrict__ w_out, double *__restrict__ f_out) {
    __shared__ float4 sp[HT];
    __shared__ double sw[HT];
    __shared__ double red[HT];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.y * s.n_atoms;
    const double *sa = sasa + (long long)blockIdx.y * s.n_typed;
    const int i = blockIdx.x * HT + tid;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    double wi = 0.0;
    if (i < nc_lim) {
        const int c = s.carbon[i];
        const float *p = coords + (base + s.typed_atom[c]) * 3;
        xi = p[0]; yi = p[1]; zi = p[2];
        const double a = sa[c];
        wi = (a > s.sa_threshold && xi <= ABSENT_X) ? tanh_area(a) : 0.0;
    }
    const bool want_field = f_out != nullptr && i < nc_lim;
    double acc = 0.0;
    for (int j0 = 0; j0 < nc_lim; j0 += HT) {
        __syncthreads();
        const int jl = j0 + tid;
        if (jl < nc_lim) {
            const int c = s.carbon[jl];
            const float *p = coords + (base + s.typed_atom[c]) * 3;
            const double a = sa[c];
            const double w = (a > s.sa_threshold && p[0] <= ABSENT_X) ? tanh_area(a) : 0.0;
            sw[tid] = w;
            sp[tid] = make_float4(p[0], p[1], p[2], w == 0.0 ? 0.f : 1.f);
        }
        __syncthreads();
        if (wi == 0.0 && !want_field) continue;
        const int n = min(HT, nc_lim - j0);
#pragma unroll 4
        for (int jj = 0; jj < n; ++jj) {
            const float4 q = sp[jj];
            if (q.w != 0.f && dist2_fast(xi, yi, zi, q.x, q.y, q.z) < HPMF_CUTOFF * HPMF_CUTOFF) {
                if (j0 + jj == i) continue;
                const double wj = sw[jj];
                acc += wj * gauss3(s, __fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z)));
            }
        }
    }
    if (i < s.n_carbon && w_out) {
        const long long o = (long long)blockIdx.y * s.n_carbon + i;
        w_out[o] = wi; f_out[o] = i < nc_lim ? acc : 0.0;
    }
    red[tid] = wi * acc;
    __syncthreads();
    for (int k = HT / 2; k > 0; k >>= 1) {
        if (tid < k) red[tid] += red[tid + k];
        __syncthreads();
    }
    if (tid == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = red[0];
}

__global__ void hpmf_finish_kernel(int n_struct, int n_tiles, const double *__restrict__ partial, double *__restrict__ v) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_struct) return;
    double sum = 0.0;
    for (int k = 0; k < n_tiles; ++k) sum += partial[(long long)c * n_tiles + k];
    v[c] = 0.5 * sum;
}


// ------------------------------------------------------------------------------------------------
// per-trial V_hpmf of (environment + one trial side chain) without re-evaluating the environment
// ------------------------------------------------------------------------------------------------
// The surface area is a product over partners, so adding the trial's atoms multiplies the environment's areas:
//   SA_k(env + trial) = SA_k(env) * prod_{n in trial, in range} term_kn ,
// and with w = tanh(SA) [SA > A_c], delta_k = w_k(env + trial) - w_k(env), F_k = sum_{j != k} w_j(env) G_kj :
//   V(env + trial) = V(env) + sum_k delta_k F_k + sum_{k < l} delta_k delta_l G_kl
//                  + sum_{n, e} w_n (w_e + delta_e) G_ne + sum_{n < m} w_n w_m G_nm
// (k, l, e: environment carbons, delta != 0 only within reach of the trial; n, m: the trial's carbons).  Exact - the
// same pair terms as the whole-structure evaluation, grouped differently.  One block per (trial, conformer).
#define MAX_NEW_T 16           // typed atoms of one side chain (ARG: 12)
#define MAX_NEW_C 12           // carbons of one side chain (TRP: 9)
#define MAX_LIST 512           // environment carbons whose weight a trial changes
#define TRIAL_MINB_DEFAULT 4

template <int MINB>
__global__ void __launch_bounds__(HT, MINB) hpmf_trial_kernel(
        HpmfDev s, const float *__restrict__ coords, const float *__restrict__ trial, const unsigned char *__restrict__ mask,
        int first_new, int n_new, int n_trials, int e1, int m, int c1, int mc, const int *__restrict__ carbon_ord,
        const double *__restrict__ sa_env, const double *__restrict__ w0, const double *__restrict__ f0,
        const double *__restrict__ v0, double *__restrict__ out, int *__restrict__ overflow,
        // inside mcsce_grow: the growth's own trial buffer (float4, per-conformer stride, row = trial * n_new + atom), the
        // residue's list of typed environment atoms that can come within SASA range of its side chain, a flag per conformer
        const float4 *__restrict__ trial4, long long trial4_stride, const int *__restrict__ near, int n_near,
        int *__restrict__ overflow_conf,
        // packed carbon positions [conformer][n_carbon] and the residue's list of carbons within Gaussian range of its side chain
        const float4 *__restrict__ cpos, const int *__restrict__ clist, int n_clist) {
    const int t = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, nt = s.n_types;
    const long long ct = (long long)c * n_trials + t;
    if (mask && !mask[ct]) { if (tid == 0) out[ct] = 0.0; return; }
    __shared__ float4 ta[MAX_NEW_T];                  // trial typed atoms: x, y, z, type
    __shared__ int tbond[MAX_NEW_T][MAXB];
    __shared__ double s_thr[MAX_SASA_TYPES * MAX_SASA_TYPES];
    __shared__ double s_R[MAX_SASA_TYPES], s_piRRs[MAX_SASA_TYPES], s_P[MAX_SASA_TYPES], s_S[MAX_SASA_TYPES];
    __shared__ double tprod[HT / 32][MAX_NEW_T];
    __shared__ float4 tc[MAX_NEW_C];                  // trial carbons
    __shared__ double tw[MAX_NEW_C];
    __shared__ float4 lpos[MAX_LIST];
    __shared__ double ldel[MAX_LIST];
    __shared__ int ln;
    __shared__ double red[HT];
    for (int k = tid; k < nt * nt; k += HT) s_thr[k] = s.thr[k];
    if (tid < nt) { s_R[tid] = s.R[tid]; s_piRRs[tid] = s.piRRs[tid]; s_P[tid] = s.P[tid]; s_S[tid] = s.S[tid]; }
    if (tid < m) {
        const int g = e1 + tid;
        if (trial4) {
            const float4 q = trial4[(long long)c * trial4_stride + (long long)t * n_new + (s.typed_atom[g] - first_new)];
            ta[tid] = make_float4(q.x, q.y, q.z, __int_as_float(s.type_of[g]));
        } else {
            const float *p = trial + (ct * n_new + (s.typed_atom[g] - first_new)) * 3;
            ta[tid] = make_float4(p[0], p[1], p[2], __int_as_float(s.type_of[g]));
        }
#pragma unroll
        for (int b = 0; b < MAXB; ++b) tbond[tid][b] = s.bonded[g * MAXB + b];
    }
    if (tid < MAX_NEW_C) { tw[tid] = 0.0; tc[tid] = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (tid == 0) ln = 0;
    __syncthreads();

    // phase 1: environment atoms x trial atoms - factors of both surface products
    double tp[MAX_NEW_T];
#pragma unroll
    for (int n = 0; n < MAX_NEW_T; ++n) tp[n] = 1.0;
    double acc = 0.0;                                  // sum_k delta_k F_k, later the Gaussian sums
    const float *cbase = coords + (long long)c * s.n_atoms * 3;
    const double *sa = sa_env + (long long)c * s.n_typed;
    const double *w0c = w0 + (long long)c * s.n_carbon, *f0c = f0 + (long long)c * s.n_carbon;
    // (with a near list only the typed atoms that can be in SASA range of the side chain are visited - a complete list,
    //  built from rigorous reach bounds, so the factors found are the same; the order of the float64 products per trial
    //  atom is the list's ascending order either way)
    // (the list of changed carbons is filled in scan order - ballots and a scan over the warps, no atomics - so that the float64
    //  sums over it do not depend on the scheduling: the same call gives the same bits)
    __shared__ int s_wcnt[HT / 32];
    const int n_scan = near ? n_near : e1;
    for (int qb = 0; qb < n_scan; qb += HT) {
        const int q0 = qb + tid;
        double del = 0.0; float xk = 0.f, yk = 0.f, zk = 0.f;
        const int k = q0 < n_scan ? (near ? near[q0] : q0) : e1;
        if (k < e1) {
            const float *p = cbase + (long long)s.typed_atom[k] * 3;
            xk = p[0]; yk = p[1]; zk = p[2];
            if (xk <= ABSENT_X) {
                int tk = -1, b0 = -1, b1 = -1, b2 = -1, b3 = -1;
                double f = 1.0;
#pragma unroll
                for (int n = 0; n < MAX_NEW_T; ++n) {
                    if (n < m) {
                        const float4 q = ta[n];
                        if (dist2_fast(xk, yk, zk, q.x, q.y, q.z) < s.pre2) {
                            const float d = __fsqrt_rn(norm2_f32(xk, yk, zk, q.x, q.y, q.z));
                            if (tk < 0) {
                                tk = s.type_of[k];
                                b0 = s.bonded[k * MAXB]; b1 = s.bonded[k * MAXB + 1]; b2 = s.bonded[k * MAXB + 2]; b3 = s.bonded[k * MAXB + 3];
                            }
                            const int tn = __float_as_int(q.w);
                            const double thr = s_thr[tk * nt + tn];
                            if ((double)d < thr) {
                                const int g = e1 + n;
                                const double pij = (g == b0 || g == b1 || g == b2 || g == b3) ? s.pij_bonded : s.pij_non_bonded;
                                f = __dmul_rn(f, sasa_term(s_P[tk], s_S[tk], s_R[tk], s_piRRs[tn], s_R[tn], thr, d, pij));
                                tp[n] = __dmul_rn(tp[n], sasa_term(s_P[tn], s_S[tn], s_R[tn], s_piRRs[tk], s_R[tk], thr, d, pij));
                            }
                        }
                    }
                }
                if (f != 1.0) {
                    const int ck = carbon_ord[k];
                    if (ck >= 0) {
                        const double a = sa[k] * f;
                        const double w = a > s.sa_threshold ? tanh_area(a) : 0.0;
                        del = w - w0c[ck];
                        if (del != 0.0) acc += del * f0c[ck];
                    }
                }
            }
        }
        const bool on = del != 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if ((tid & 31) == 0) s_wcnt[tid >> 5] = __popc(bal);
        __syncthreads();
        int base = ln;
        for (int w = 0; w < (tid >> 5); ++w) base += s_wcnt[w];
        if (on) {
            const int idx = base + __popc(bal & ((1u << (tid & 31)) - 1u));
            if (idx < MAX_LIST) { lpos[idx] = make_float4(xk, yk, zk, 0.f); ldel[idx] = del; }
        }
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < HT / 32; ++w) tot += s_wcnt[w]; ln += tot; }
        __syncthreads();
    }
    // block-wide products for the trial atoms
#pragma unroll
    for (int n = 0; n < MAX_NEW_T; ++n) {
        if (n < m) {
            double v = tp[n];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, off);
            if ((tid & 31) == 0) tprod[tid >> 5][n] = v;
        }
    }
    __syncthreads();
    // phase 2: trial atoms among themselves -> their surface areas and weights
    if (tid < m) {
        double prod = 1.0;
#pragma unroll
        for (int w = 0; w < HT / 32; ++w) prod *= tprod[w][tid];
        const float4 me = ta[tid];
        const int ti = __float_as_int(me.w);
        for (int n = 0; n < m; ++n) {
            if (n == tid) continue;
            const float4 q = ta[n];
            const float d = __fsqrt_rn(norm2_f32(me.x, me.y, me.z, q.x, q.y, q.z));
            const int tn = __float_as_int(q.w);
            const double thr = s_thr[ti * nt + tn];
            if ((double)d < thr) {
                const int g = e1 + n;
                const bool bonded = g == tbond[tid][0] || g == tbond[tid][1] || g == tbond[tid][2] || g == tbond[tid][3];
                prod = __dmul_rn(prod, sasa_term(s_P[ti], s_S[ti], s_R[ti], s_piRRs[tn], s_R[tn], thr, d,
                                                 bonded ? s.pij_bonded : s.pij_non_bonded));
            }
        }
        const int ck = carbon_ord[e1 + tid];
        if (ck >= 0) {
            const double a = __dmul_rn(s_S[ti], prod);
            tw[ck - c1] = a > s.sa_threshold ? tanh_area(a) : 0.0;
            tc[ck - c1] = make_float4(me.x, me.y, me.z, 0.f);
        }
    }
    __syncthreads();
    const int L = min(ln, MAX_LIST);
    if (tid == 0 && ln > MAX_LIST) { atomicExch(overflow, 1); if (overflow_conf) overflow_conf[c] = 1; }
    const float cut2 = HPMF_CUTOFF * HPMF_CUTOFF;
    // phase 3a: trial carbons x environment carbons (weights of the environment alone)
    {
        const int n3 = clist ? n_clist : c1;
        const float4 *cposc = cpos ? cpos + (long long)c * s.n_carbon : nullptr;
        for (int q0 = tid; q0 < n3; q0 += HT) {
            const int ck = clist ? clist[q0] : q0;
            if (ck >= c1) continue;
            const double wk = w0c[ck];
            if (wk == 0.0) continue;
            float xk, yk, zk;
            if (cposc) { const float4 p = cposc[ck]; xk = p.x; yk = p.y; zk = p.z; }
            else { const float *p = cbase + (long long)s.typed_atom[s.carbon[ck]] * 3; xk = p[0]; yk = p[1]; zk = p[2]; }
            for (int n = 0; n < mc; ++n) {
                const double wn = tw[n];
                const float4 q = tc[n];
                if (wn != 0.0 && dist2_fast(xk, yk, zk, q.x, q.y, q.z) < cut2)
                    acc += wn * wk * gauss3(s, __fsqrt_rn(norm2_f32(xk, yk, zk, q.x, q.y, q.z)));
            }
        }
    }
    // phase 3b: the changed environment carbons: with the trial carbons, and among themselves
    for (int a = tid; a < L; a += HT) {
        const float4 pa = lpos[a];
        const double da = ldel[a];
        for (int n = 0; n < mc; ++n) {
            const double wn = tw[n];
            const float4 q = tc[n];
            if (wn != 0.0 && dist2_fast(pa.x, pa.y, pa.z, q.x, q.y, q.z) < cut2)
                acc += wn * da * gauss3(s, __fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, q.x, q.y, q.z)));
        }
        for (int b = a + 1; b < L; ++b) {
            const float4 pb = lpos[b];
            if (dist2_fast(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < cut2)
                acc += da * ldel[b] * gauss3(s, __fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z)));
        }
    }
    // phase 3c: trial carbons among themselves
    if (tid < mc) {
        const float4 pa = tc[tid];
        for (int b = tid + 1; b < mc; ++b) {
            const float4 pb = tc[b];
            if (tw[tid] != 0.0 && tw[b] != 0.0 && dist2_fast(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < cut2)
                acc += tw[tid] * tw[b] * gauss3(s, __fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z)));
        }
    }
    red[tid] = acc;
    __syncthreads();
    for (int k = HT / 2; k > 0; k >>= 1) {
        if (tid < k) red[tid] += red[tid + k];
        __syncthreads();
    }
    if (tid == 0) out[ct] = v0[c] + red[0];
}


// ------------------------------------------------------------------------------------------------
// the same per-trial evaluation, one block per CONFORMER and one warp per trial (the growth's kernel)
// ------------------------------------------------------------------------------------------------
// hpmf_trial_kernel spends a block per (trial, conformer): every block reloads the type tables, gathers the environment atoms
// through two index tables and evaluates the float64 surface factors under heavy divergence (a few lanes of every warp are in
// range) - 262 us per residue at 256 conformers of config 3, 14 us per block for ~3000 instructions of work.  Here a block
// takes a conformer: the residue's environment lists are staged in shared memory once (typed atoms in surface range with their
// areas; the EXPOSED carbons in Gaussian range with their weights, compacted in order), then each warp takes trials in turn:
//   pass A  lanes over the staged atoms, float32 range test against the trial's <= 16 typed atoms -> (atom, bit mask) hit list
//   pass B  lanes over the hit list (dense: every lane has work): exact distances, float64 factors for both surface products
//   phase 2 the trial's own atoms, lane = atom;  phase 3 Gaussian sums, lanes over the staged carbons / the changed carbons
// Same terms in the same arithmetic as hpmf_trial_kernel (float64 products are taken in another order: 1e-13 relative).
// Measured alternatives (round 2, config 3, ms per 2048 conformers with the term / without = 540): this kernel 893; a version
// with a bounding-sphere prefilter, one (atom, trial atom) pair per lane in pass B and 16 warps per block 995 (8 warps 1185):
// the denser lanes did not pay for the extra list passes.
#define CT_WARPS 8
#define CT_THREADS (CT_WARPS * 32)
#define CT_TRIALS 1024         // trials per residue the kernel indexes (the reference's largest set: 729)
#define CT_NEAR 896            // staged typed atoms in surface range of the residue (more: the conformer is given up, ok = 2)
#define CT_CARB 640            // staged exposed carbons in Gaussian range (more: global reads)
#define CT_HITS 192            // (atom, mask) pairs one trial can hit
#define CT_LIST 160            // environment carbons whose weight one trial changes

struct ConfTrialArgs {
    HpmfDev s;
    const float *coords; const float4 *trial4; long long trial4_stride; const unsigned char *mask;
    int first_new, n_new, n_trials, e1, m, c1, mc;
    const int *carbon_ord; const double *sa_env, *w0, *f0, *v0; double *out;
    int *overflow, *overflow_conf;
    const int *near; int n_near; const int *clist; int n_clist;
    const float4 *cpos;
    float4 *scratch_pos; double *scratch_w;          // [conformer][n_carbon] exposed carbons beyond the shared-memory stage
    // The lists come from bounds that hold for every conformer; with the conformer's own coordinates they shrink to the atoms
    // within reach of the residue: centre = its CA, r2_near = (reach + surface range)^2, r2_carb = (reach + Gaussian cutoff)^2
    // (r2 <= 0: no filter).
    float cax, cay, caz, r2_near, r2_carb;
};

struct ConfTrialSmem {
    double thr[MAX_SASA_TYPES * MAX_SASA_TYPES];
    double R[MAX_SASA_TYPES], piRRs[MAX_SASA_TYPES], P[MAX_SASA_TYPES], S[MAX_SASA_TYPES], invS[MAX_SASA_TYPES];
    float4 near_pos[CT_NEAR];      // x, y, z, type (int bits): the typed atoms in surface range of the residue, compacted in order
    double near_sa[CT_NEAR];
    int near_k[CT_NEAR];           // their typed index
    int n_near;
    float4 carb_pos[CT_CARB];
    double carb_w[CT_CARB];
    int trials[CT_TRIALS];         // the conformer's trials that passed the clash filter, ascending
    int n_active, n_carb;
    int warp_cnt[CT_WARPS];
    // per warp
    float4 ta[CT_WARPS][MAX_NEW_T];
    float4 tc[CT_WARPS][MAX_NEW_C];
    double tw[CT_WARPS][MAX_NEW_C];
    int hit_q[CT_WARPS][CT_HITS];
    unsigned hit_m[CT_WARPS][CT_HITS];
    float4 lpos[CT_WARPS][CT_LIST];
    double ldel[CT_WARPS][CT_LIST];
};

__device__ __forceinline__ double warp_sum_fixed(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__global__ void __launch_bounds__(CT_THREADS, 2) hpmf_trials_conf_kernel(ConfTrialArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ConfTrialSmem &sm = *reinterpret_cast<ConfTrialSmem *>(smem_raw);
    const HpmfDev &s = a.s;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = s.n_types;
    const int T = a.n_trials, e1 = a.e1, m = a.m, c1 = a.c1, mc = a.mc;
    const unsigned char *mask = a.mask ? a.mask + (long long)c * T : nullptr;
    double *out = a.out + (long long)c * T;
    const float *cbase = a.coords + (long long)c * s.n_atoms * 3;
    const double *sa = a.sa_env + (long long)c * s.n_typed;
    const double *w0c = a.w0 + (long long)c * s.n_carbon, *f0c = a.f0 + (long long)c * s.n_carbon;
    const float cut2 = HPMF_CUTOFF * HPMF_CUTOFF;

    // ---- the conformer's active trials (ordered compaction of the mask), zeros for the others ----
    if (tid == 0) { sm.n_active = 0; sm.n_carb = 0; }
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += CT_THREADS) {
        const int t = t0 + tid;
        const bool on = t < T && (!mask || mask[t]);
        if (t < T && !on) out[t] = 0.0;
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (lane == 0) sm.warp_cnt[warp] = __popc(b);
        __syncthreads();
        int base = sm.n_active;
        for (int w = 0; w < warp; ++w) base += sm.warp_cnt[w];
        if (on) sm.trials[base + __popc(b & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < CT_WARPS; ++w) tot += sm.warp_cnt[w]; sm.n_active += tot; }
        __syncthreads();
    }
    const int n_active = sm.n_active;
    if (n_active == 0) return;

    // ---- tables and the residue's environment lists, once per conformer ----
    for (int k = tid; k < nt * nt; k += CT_THREADS) sm.thr[k] = s.thr[k];
    if (tid < nt) {
        sm.R[tid] = s.R[tid]; sm.piRRs[tid] = s.piRRs[tid]; sm.P[tid] = s.P[tid]; sm.S[tid] = s.S[tid];
        sm.invS[tid] = recip_f64(s.S[tid]);
    }
    {   // typed atoms in surface range: present, and (with the conformer's own coordinates) within reach of the residue
        const int n0 = a.near ? a.n_near : e1;
        if (tid == 0) sm.n_near = 0;
        __syncthreads();
        for (int q0 = 0; q0 < n0; q0 += CT_THREADS) {
            const int q = q0 + tid;
            int k = -1; float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            bool on = false;
            if (q < n0) {
                k = a.near ? a.near[q] : q;
                if (k < e1) {
                    const float *p = cbase + (long long)s.typed_atom[k] * 3;
                    v = make_float4(p[0], p[1], p[2], 0.f);
                    on = v.x <= ABSENT_X && (a.r2_near <= 0.f || dist2_fast(v.x, v.y, v.z, a.cax, a.cay, a.caz) < a.r2_near);
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, on);
            if (lane == 0) sm.warp_cnt[warp] = __popc(b);
            __syncthreads();
            int base = sm.n_near;
            for (int ww = 0; ww < warp; ++ww) base += sm.warp_cnt[ww];
            if (on) {
                const int idx = base + __popc(b & ((1u << lane) - 1u));
                if (idx < CT_NEAR) {
                    v.w = __int_as_float(s.type_of[k]);
                    sm.near_pos[idx] = v; sm.near_sa[idx] = sa[k]; sm.near_k[idx] = k;
                }
            }
            __syncthreads();
            if (tid == 0) { int tot = 0; for (int ww = 0; ww < CT_WARPS; ++ww) tot += sm.warp_cnt[ww]; sm.n_near += tot; }
            __syncthreads();
        }
    }
    const int n_near = min(sm.n_near, CT_NEAR);
    const bool near_over = sm.n_near > CT_NEAR;
    {   // exposed carbons in Gaussian range, in list order
        const int n3 = a.clist ? a.n_clist : c1;
        const float4 *cposc = a.cpos + (long long)c * s.n_carbon;
        float4 *spos = a.scratch_pos + (long long)c * s.n_carbon;
        double *sw = a.scratch_w + (long long)c * s.n_carbon;
        for (int q0 = 0; q0 < n3; q0 += CT_THREADS) {
            const int q = q0 + tid;
            int ck = -1; double w = 0.0;
            float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < n3) {
                ck = a.clist ? a.clist[q] : q;
                if (ck < c1) {
                    w = w0c[ck];
                    if (w != 0.0) { pc = cposc[ck]; if (a.r2_carb > 0.f && !(dist2_fast(pc.x, pc.y, pc.z, a.cax, a.cay, a.caz) < a.r2_carb)) w = 0.0; }
                }
            }
            const bool on = w != 0.0;
            const unsigned b = __ballot_sync(0xffffffffu, on);
            if (lane == 0) sm.warp_cnt[warp] = __popc(b);
            __syncthreads();
            int base = sm.n_carb;
            for (int ww = 0; ww < warp; ++ww) base += sm.warp_cnt[ww];
            if (on) {
                const int idx = base + __popc(b & ((1u << lane) - 1u));
                if (idx < CT_CARB) { sm.carb_pos[idx] = pc; sm.carb_w[idx] = w; }
                else { spos[idx] = pc; sw[idx] = w; }
            }
            __syncthreads();
            if (tid == 0) { int tot = 0; for (int ww = 0; ww < CT_WARPS; ++ww) tot += sm.warp_cnt[ww]; sm.n_carb += tot; }
            __syncthreads();
        }
    }
    __syncthreads();
    const int n_carb = sm.n_carb;
    const float4 *spos = a.scratch_pos + (long long)c * s.n_carbon;
    const double *sw = a.scratch_w + (long long)c * s.n_carbon;

    // ---- trials, one per warp at a time ----
    float4 *ta = sm.ta[warp]; float4 *tc = sm.tc[warp]; double *tw = sm.tw[warp];
    int *hit_q = sm.hit_q[warp]; unsigned *hit_m = sm.hit_m[warp];
    float4 *lpos = sm.lpos[warp]; double *ldel = sm.ldel[warp];
    for (int it = warp; it < n_active; it += CT_WARPS) {
        const int t = sm.trials[it];
        __syncwarp();
        int tb0 = -1, tb1 = -1, tb2 = -1, tb3 = -1;
        if (lane < m) {
            const int g = e1 + lane;
            const float4 q = a.trial4[(long long)c * a.trial4_stride + (long long)t * a.n_new + (s.typed_atom[g] - a.first_new)];
            ta[lane] = make_float4(q.x, q.y, q.z, __int_as_float(s.type_of[g]));
            tb0 = s.bonded[g * MAXB]; tb1 = s.bonded[g * MAXB + 1]; tb2 = s.bonded[g * MAXB + 2]; tb3 = s.bonded[g * MAXB + 3];
        }
        if (lane < MAX_NEW_C) { tw[lane] = 0.0; tc[lane] = make_float4(0.f, 0.f, 0.f, 0.f); }
        __syncwarp();
        // pass A: which staged atoms have a trial atom within the (padded) surface range.  The list is in sequence order, so
        // the 32 atoms of an iteration are neighbours in space: a test against the trial's bounding sphere lets whole
        // iterations skip the per-atom tests.
        float bx, by, bz, brad2;
        {
            float sx = lane < m ? ta[lane].x : 0.f, sy = lane < m ? ta[lane].y : 0.f, sz = lane < m ? ta[lane].z : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, off); sy += __shfl_xor_sync(0xffffffffu, sy, off); sz += __shfl_xor_sync(0xffffffffu, sz, off);
            }
            bx = sx / (float)m; by = sy / (float)m; bz = sz / (float)m;
            float r = lane < m ? sqrtf(dist2_fast(bx, by, bz, ta[lane].x, ta[lane].y, ta[lane].z)) : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, off));
            const float brad = r * 1.0001f + sqrtf(s.pre2) + 1.0e-3f;
            brad2 = brad * brad;
        }
        int n_hits = 0;
        for (int q0 = 0; q0 < n_near; q0 += 32) {
            const int q = q0 + lane;
            unsigned bits = 0;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            bool in = false;
            if (q < n_near) { p = sm.near_pos[q]; in = dist2_fast(p.x, p.y, p.z, bx, by, bz) < brad2; }
            if (!__any_sync(0xffffffffu, in)) continue;
            if (in) {
#pragma unroll
                for (int n = 0; n < MAX_NEW_T; ++n)
                    if (n < m) { const float4 u = ta[n]; if (dist2_fast(p.x, p.y, p.z, u.x, u.y, u.z) < s.pre2) bits |= 1u << n; }
            }
            const unsigned b = __ballot_sync(0xffffffffu, bits != 0);
            if (bits != 0) {
                const int idx = n_hits + __popc(b & ((1u << lane) - 1u));
                if (idx < CT_HITS) { hit_q[idx] = q; hit_m[idx] = bits; }
            }
            n_hits += __popc(b);
        }
        bool over = n_hits > CT_HITS || near_over;
        n_hits = min(n_hits, CT_HITS);
        __syncwarp();
        // pass B: the factors of both surface products, lanes over the hit list
        double tp[MAX_NEW_T];
#pragma unroll
        for (int n = 0; n < MAX_NEW_T; ++n) tp[n] = 1.0;
        double acc = 0.0;
        int n_list = 0;
        for (int h0 = 0; h0 < n_hits; h0 += 32) {
            const int hh = h0 + lane;
            const bool have = hh < n_hits;
            double f = 1.0;
            int k = -1; float xk = 0.f, yk = 0.f, zk = 0.f; double area = 0.0;
            if (have) {
                const int q = hit_q[hh];
                const unsigned bits = hit_m[hh];
                k = sm.near_k[q];
                int tk;
                { const float4 p = sm.near_pos[q]; xk = p.x; yk = p.y; zk = p.z; tk = __float_as_int(p.w); area = sm.near_sa[q]; }
                const int4 bk = *reinterpret_cast<const int4 *>(s.bonded + (long long)k * MAXB);
#pragma unroll
                for (int n = 0; n < MAX_NEW_T; ++n) {
                    if ((bits >> n) & 1u) {
                        const float4 u = ta[n];
                        const float d = __fsqrt_rn(norm2_f32(xk, yk, zk, u.x, u.y, u.z));
                        const int tn = __float_as_int(u.w);
                        const double thr = sm.thr[tk * nt + tn];
                        if ((double)d < thr) {
                            const int g = e1 + n;
                            const double pij = (g == bk.x || g == bk.y || g == bk.z || g == bk.w) ? s.pij_bonded : s.pij_non_bonded;
                            double t_kn, t_nk;
                            sasa_term_pair(sm.P[tk], sm.invS[tk], sm.R[tk], sm.piRRs[tk], sm.P[tn], sm.invS[tn], sm.R[tn], sm.piRRs[tn],
                                           thr, d, pij, t_kn, t_nk);
                            f *= t_kn; tp[n] *= t_nk;
                        }
                    }
                }
            }
            double del = 0.0;
            if (have && f != 1.0) {
                const int ck = a.carbon_ord[k];
                if (ck >= 0) {
                    const double ar = area * f;
                    const double w = ar > s.sa_threshold ? tanh_area(ar) : 0.0;
                    del = w - w0c[ck];
                    if (del != 0.0) acc += del * f0c[ck];
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, del != 0.0);
            if (del != 0.0) {
                const int idx = n_list + __popc(b & ((1u << lane) - 1u));
                if (idx < CT_LIST) { lpos[idx] = make_float4(xk, yk, zk, 0.f); ldel[idx] = del; }
            }
            n_list += __popc(b);
        }
        over |= n_list > CT_LIST;
        const int L = min(n_list, CT_LIST);
        // products over the lanes, per trial atom (fixed butterfly order)
#pragma unroll
        for (int n = 0; n < MAX_NEW_T; ++n) {
            if (n < m) {
                double v = tp[n];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, off);
                tp[n] = v;
            }
        }
        // phase 2: the trial's atoms among themselves -> their areas and weights (lane = trial atom)
        {
            double prod = 1.0;
#pragma unroll
            for (int n = 0; n < MAX_NEW_T; ++n) if (n == lane) prod = tp[n];
            if (lane < m) {
                const float4 me = ta[lane];
                const int ti = __float_as_int(me.w);
                for (int n = 0; n < m; ++n) {
                    if (n == lane) continue;
                    const float4 q = ta[n];
                    const float d = __fsqrt_rn(norm2_f32(me.x, me.y, me.z, q.x, q.y, q.z));
                    const int tn = __float_as_int(q.w);
                    const double thr = sm.thr[ti * nt + tn];
                    if ((double)d < thr) {
                        const int g = e1 + n;
                        const bool bonded = g == tb0 || g == tb1 || g == tb2 || g == tb3;
                        prod = __dmul_rn(prod, sasa_term(sm.P[ti], sm.S[ti], sm.R[ti], sm.piRRs[tn], sm.R[tn], thr, d,
                                                         bonded ? s.pij_bonded : s.pij_non_bonded));
                    }
                }
                const int ck = a.carbon_ord[e1 + lane];
                if (ck >= 0) {
                    const double ar = __dmul_rn(sm.S[ti], prod);
                    tw[ck - c1] = ar > s.sa_threshold ? tanh_area(ar) : 0.0;
                    tc[ck - c1] = make_float4(me.x, me.y, me.z, 0.f);
                }
            }
        }
        __syncwarp();
        // phase 3a: trial carbons x exposed environment carbons (weights of the environment alone)
        for (int q = lane; q < n_carb; q += 32) {
            float4 p; double wk;
            if (q < CT_CARB) { p = sm.carb_pos[q]; wk = sm.carb_w[q]; } else { p = spos[q]; wk = sw[q]; }
            for (int n = 0; n < mc; ++n) {
                const double wn = tw[n];
                const float4 u = tc[n];
                if (wn != 0.0 && dist2_fast(p.x, p.y, p.z, u.x, u.y, u.z) < cut2)
                    acc += wn * wk * gauss3(s, __fsqrt_rn(norm2_f32(p.x, p.y, p.z, u.x, u.y, u.z)));
            }
        }
        // phase 3b: the changed environment carbons: with the trial carbons, and among themselves
        for (int i = lane; i < L; i += 32) {
            const float4 pa = lpos[i];
            const double da = ldel[i];
            for (int n = 0; n < mc; ++n) {
                const double wn = tw[n];
                const float4 u = tc[n];
                if (wn != 0.0 && dist2_fast(pa.x, pa.y, pa.z, u.x, u.y, u.z) < cut2)
                    acc += wn * da * gauss3(s, __fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, u.x, u.y, u.z)));
            }
            for (int j = i + 1; j < L; ++j) {
                const float4 pb = lpos[j];
                if (dist2_fast(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < cut2)
                    acc += da * ldel[j] * gauss3(s, __fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z)));
            }
        }
        // phase 3c: trial carbons among themselves
        if (lane < mc) {
            const float4 pa = tc[lane];
            for (int j = lane + 1; j < mc; ++j) {
                const float4 pb = tc[j];
                if (tw[lane] != 0.0 && tw[j] != 0.0 && dist2_fast(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < cut2)
                    acc += tw[lane] * tw[j] * gauss3(s, __fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z)));
            }
        }
        acc = warp_sum_fixed(acc);
        if (lane == 0) {
            out[t] = a.v0[c] + acc;
            if (over) { atomicExch(a.overflow, 1); if (a.overflow_conf) a.overflow_conf[c] = 1; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// carrying the environment state from one growth step to the next
// ------------------------------------------------------------------------------------------------
// Between two steps the environment only gains the atoms [L, F): the side chain selected at the previous step (and
// conformer-independent GLY/ALA side chains in between).  Instead of re-deriving areas, weights and fields of the
// whole environment, the state of level L is extended: the new atoms multiply the areas of the atoms in their reach
// (hpmf_extend_kernel, one block per conformer - the trial kernel's first two phases with the results written
// back), then every carbon's field gains the changed weights' and the new carbons' terms and the new carbons get
// their full sums (hpmf_field_update_kernel).
#define EXT_LIST 1024          // environment carbons whose weight one extension changes

__global__ void __launch_bounds__(HT) hpmf_extend_kernel(
        HpmfDev s, const float *__restrict__ coords, int eL, int m, const int *__restrict__ carbon_ord,
        double *__restrict__ sa_env, double *__restrict__ w0, float4 *__restrict__ ext_pos, double *__restrict__ ext_del,
        int *__restrict__ ext_ck, int *__restrict__ ext_n, int *__restrict__ overflow, int *__restrict__ overflow_conf) {
    const int c = blockIdx.x, tid = threadIdx.x, nt = s.n_types;
    __shared__ float4 ta[MAX_NEW_T];
    __shared__ int tbond[MAX_NEW_T][MAXB];
    __shared__ double s_thr[MAX_SASA_TYPES * MAX_SASA_TYPES];
    __shared__ double s_R[MAX_SASA_TYPES], s_piRRs[MAX_SASA_TYPES], s_P[MAX_SASA_TYPES], s_S[MAX_SASA_TYPES];
    __shared__ double tprod[HT / 32][MAX_NEW_T];
    __shared__ int ln;
    const float *cbase = coords + (long long)c * s.n_atoms * 3;
    double *sa = sa_env + (long long)c * s.n_typed;
    double *w0c = w0 + (long long)c * s.n_carbon;
    for (int k = tid; k < nt * nt; k += HT) s_thr[k] = s.thr[k];
    if (tid < nt) { s_R[tid] = s.R[tid]; s_piRRs[tid] = s.piRRs[tid]; s_P[tid] = s.P[tid]; s_S[tid] = s.S[tid]; }
    if (tid < m) {
        const int g = eL + tid;
        const float *p = cbase + (long long)s.typed_atom[g] * 3;
        ta[tid] = make_float4(p[0], p[1], p[2], __int_as_float(s.type_of[g]));     // an absent atom (1e15) is out of everyone's reach
#pragma unroll
        for (int b = 0; b < MAXB; ++b) tbond[tid][b] = s.bonded[g * MAXB + b];
    }
    if (tid == 0) ln = 0;
    __syncthreads();
    // bounding sphere of the new atoms (padded by the surface range): atoms outside it skip the per-atom tests, and as the
    // environment is walked in sequence order whole warps do
    __shared__ float s_sph[4];
    if (tid == 0) {
        float cx = 0.f, cy = 0.f, cz = 0.f; int cnt = 0;
        for (int n = 0; n < m; ++n) if (ta[n].x <= ABSENT_X) { cx += ta[n].x; cy += ta[n].y; cz += ta[n].z; ++cnt; }
        if (cnt) { cx /= cnt; cy /= cnt; cz /= cnt; }
        float r = 0.f;
        for (int n = 0; n < m; ++n) if (ta[n].x <= ABSENT_X) r = fmaxf(r, sqrtf(dist2_fast(cx, cy, cz, ta[n].x, ta[n].y, ta[n].z)));
        r = cnt ? r * 1.0001f + sqrtf(s.pre2) + 1.0e-3f : -1.f;
        s_sph[0] = cx; s_sph[1] = cy; s_sph[2] = cz; s_sph[3] = r < 0.f ? -1.f : r * r;
    }
    __syncthreads();
    const float sx = s_sph[0], sy = s_sph[1], sz = s_sph[2], sr2 = s_sph[3];
    double tp[MAX_NEW_T];
#pragma unroll
    for (int n = 0; n < MAX_NEW_T; ++n) tp[n] = 1.0;
    // (the list of changed carbons is filled in atom order - ballots and a scan over the warps, no atomics - so that the
    //  float64 sums taken over it later do not depend on the scheduling: repeated runs are bitwise identical)
    __shared__ int s_wcnt[HT / 32];
    for (int k0 = 0; k0 < eL; k0 += HT) {
        const int k = k0 + tid;
        double del = 0.0; float xk = 0.f, yk = 0.f, zk = 0.f; int ck = -1;
        if (k < eL) {
            const float *p = cbase + (long long)s.typed_atom[k] * 3;
            xk = p[0]; yk = p[1]; zk = p[2];
            if (xk <= ABSENT_X && dist2_fast(xk, yk, zk, sx, sy, sz) < sr2) {
                int tk = -1, b0 = -1, b1 = -1, b2 = -1, b3 = -1;
                double f = 1.0;
#pragma unroll
                for (int n = 0; n < MAX_NEW_T; ++n) {
                    if (n < m) {
                        const float4 q = ta[n];
                        if (dist2_fast(xk, yk, zk, q.x, q.y, q.z) < s.pre2) {
                            const float d = __fsqrt_rn(norm2_f32(xk, yk, zk, q.x, q.y, q.z));
                            if (tk < 0) {
                                tk = s.type_of[k];
                                b0 = s.bonded[k * MAXB]; b1 = s.bonded[k * MAXB + 1]; b2 = s.bonded[k * MAXB + 2]; b3 = s.bonded[k * MAXB + 3];
                            }
                            const int tn = __float_as_int(q.w);
                            const double thr = s_thr[tk * nt + tn];
                            if ((double)d < thr) {
                                const int g = eL + n;
                                const double pij = (g == b0 || g == b1 || g == b2 || g == b3) ? s.pij_bonded : s.pij_non_bonded;
                                f = __dmul_rn(f, sasa_term(s_P[tk], s_S[tk], s_R[tk], s_piRRs[tn], s_R[tn], thr, d, pij));
                                tp[n] = __dmul_rn(tp[n], sasa_term(s_P[tn], s_S[tn], s_R[tn], s_piRRs[tk], s_R[tk], thr, d, pij));
                            }
                        }
                    }
                }
                if (f != 1.0) {
                    const double a = sa[k] * f;
                    sa[k] = a;
                    ck = carbon_ord[k];
                    if (ck >= 0) {
                        const double w = a > s.sa_threshold ? tanh_area(a) : 0.0;
                        del = w - w0c[ck];
                        if (del != 0.0) w0c[ck] = w;
                    }
                }
            }
        }
        const bool on = del != 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if ((tid & 31) == 0) s_wcnt[tid >> 5] = __popc(bal);
        __syncthreads();
        int base = ln;
        for (int w = 0; w < (tid >> 5); ++w) base += s_wcnt[w];
        if (on) {
            const int idx = base + __popc(bal & ((1u << (tid & 31)) - 1u));
            if (idx < EXT_LIST) {
                const long long o = (long long)c * EXT_LIST + idx;
                ext_pos[o] = make_float4(xk, yk, zk, 0.f); ext_del[o] = del; ext_ck[o] = ck;
            }
        }
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < HT / 32; ++w) tot += s_wcnt[w]; ln += tot; }
        __syncthreads();
    }
#pragma unroll
    for (int n = 0; n < MAX_NEW_T; ++n) {
        if (n < m) {
            double v = tp[n];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, off);
            if ((tid & 31) == 0) tprod[tid >> 5][n] = v;
        }
    }
    __syncthreads();
    if (tid < m) {
        double prod = 1.0;
#pragma unroll
        for (int w = 0; w < HT / 32; ++w) prod *= tprod[w][tid];
        const float4 me = ta[tid];
        const int ti = __float_as_int(me.w);
        const bool absent = me.x > ABSENT_X;
        for (int n = 0; n < m && !absent; ++n) {
            if (n == tid) continue;
            const float4 q = ta[n];
            if (q.x > ABSENT_X) continue;
            const float d = __fsqrt_rn(norm2_f32(me.x, me.y, me.z, q.x, q.y, q.z));
            const int tn = __float_as_int(q.w);
            const double thr = s_thr[ti * nt + tn];
            if ((double)d < thr) {
                const int g = eL + n;
                const bool bonded = g == tbond[tid][0] || g == tbond[tid][1] || g == tbond[tid][2] || g == tbond[tid][3];
                prod = __dmul_rn(prod, sasa_term(s_P[ti], s_S[ti], s_R[ti], s_piRRs[tn], s_R[tn], thr, d,
                                                 bonded ? s.pij_bonded : s.pij_non_bonded));
            }
        }
        const double a = absent ? 0.0 : __dmul_rn(s_S[ti], prod);
        sa[eL + tid] = a;
        const int ck = carbon_ord[eL + tid];
        if (ck >= 0) w0c[ck] = a > s.sa_threshold ? tanh_area(a) : 0.0;
    }
    if (tid == 0) {
        ext_n[c] = min(ln, EXT_LIST);
        if (ln > EXT_LIST) { atomicExch(overflow, 1); if (overflow_conf) overflow_conf[c] = 1; }
    }
}

// fields and V of the extended environment.  Kernel 1 (tiles over the carbons [0, cL) that were there before): every carbon's
// field gains the changed weights' terms (list staged in shared memory) and the new carbons' terms; the same Gaussians, weighted
// by the old carbon, are the new carbons' fields - summed per tile in a fixed order (shuffles, then the four warps), so that no
// thread walks all carbons alone (the first version had the <= 12 new carbons do that: 225 us per step at 256 conformers of
// config 3, 2/3 of it waiting for those threads).  Kernel 2 (one warp per conformer): the new carbons' fields from the tile sums
// plus their mutual terms, then V = 1/2 sum w F.
#define EXT_NEW_C MAX_NEW_T

__global__ void __launch_bounds__(HT) hpmf_field_update_kernel(
        HpmfDev s, const float *__restrict__ coords, int cL, int cF, const double *__restrict__ w0, double *__restrict__ f0,
        const float4 *__restrict__ ext_pos, const double *__restrict__ ext_del, const int *__restrict__ ext_ck,
        const int *__restrict__ ext_n, double *__restrict__ partial, double *__restrict__ partial_new) {
    __shared__ float4 s_pos[EXT_LIST];
    __shared__ double s_del[EXT_LIST];
    __shared__ int s_ck[EXT_LIST];
    __shared__ float4 s_np[EXT_NEW_C];
    __shared__ double s_nw[EXT_NEW_C];
    __shared__ double s_red[HT / 32][EXT_NEW_C + 1];
    const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * HT + tid;
    const float *cbase = coords + (long long)c * s.n_atoms * 3;
    const double *w0c = w0 + (long long)c * s.n_carbon;
    double *f0c = f0 + (long long)c * s.n_carbon;
    const float cut2 = HPMF_CUTOFF * HPMF_CUTOFF;
    const int L = ext_n[c], nn = min(cF - cL, EXT_NEW_C);
    for (int l = tid; l < L; l += HT) {
        const long long o = (long long)c * EXT_LIST + l;
        s_pos[l] = ext_pos[o]; s_del[l] = ext_del[o]; s_ck[l] = ext_ck[o];
    }
    if (tid < nn) {
        const float *q = cbase + (long long)s.typed_atom[s.carbon[cL + tid]] * 3;
        s_np[tid] = make_float4(q[0], q[1], q[2], 0.f);
        s_nw[tid] = q[0] > ABSENT_X ? 0.0 : w0c[cL + tid];
    }
    __syncthreads();
    double contrib = 0.0, gn[EXT_NEW_C];
#pragma unroll
    for (int n = 0; n < EXT_NEW_C; ++n) gn[n] = 0.0;
    if (i < cL) {
        const float *p = cbase + (long long)s.typed_atom[s.carbon[i]] * 3;
        const float xi = p[0], yi = p[1], zi = p[2];
        const double wi = w0c[i];
        double f = 0.0;
        if (xi <= ABSENT_X) {
            f = f0c[i];
            for (int l = 0; l < L; ++l) {
                if (s_ck[l] == i) continue;
                const float4 q = s_pos[l];
                if (dist2_fast(xi, yi, zi, q.x, q.y, q.z) < cut2)
                    f += s_del[l] * gauss3(s, __fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z)));
            }
#pragma unroll
            for (int n = 0; n < EXT_NEW_C; ++n) {
                if (n < nn) {
                    const float4 q = s_np[n];
                    if (q.x <= ABSENT_X && dist2_fast(xi, yi, zi, q.x, q.y, q.z) < cut2) {
                        const double g = gauss3(s, __fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z)));
                        f += s_nw[n] * g;
                        gn[n] = wi * g;
                    }
                }
            }
        }
        f0c[i] = f;
        contrib = wi * f;
    }
    // fixed-order sums: lanes by shuffles, then the warps one after the other
#pragma unroll
    for (int n = 0; n <= EXT_NEW_C; ++n) {
        if (n < nn || n == EXT_NEW_C) {
            double v = n == EXT_NEW_C ? contrib : gn[n];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) s_red[warp][n] = v;
        }
    }
    __syncthreads();
    if (tid <= EXT_NEW_C && (tid < nn || tid == EXT_NEW_C)) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < HT / 32; ++w) v += s_red[w][tid];
        const long long t = (long long)c * gridDim.x + blockIdx.x;
        if (tid == EXT_NEW_C) partial[t] = v; else partial_new[t * EXT_NEW_C + tid] = v;
    }
}

__global__ void __launch_bounds__(32) hpmf_field_finish_kernel(
        HpmfDev s, const float *__restrict__ coords, int cL, int cF, int n_tiles, const double *__restrict__ w0,
        double *__restrict__ f0, const double *__restrict__ partial, const double *__restrict__ partial_new, double *__restrict__ v0) {
    __shared__ float4 s_np[EXT_NEW_C];
    __shared__ double s_nw[EXT_NEW_C];
    const int c = blockIdx.x, tid = threadIdx.x;
    const float *cbase = coords + (long long)c * s.n_atoms * 3;
    const double *w0c = w0 + (long long)c * s.n_carbon;
    const int nn = min(cF - cL, EXT_NEW_C);
    if (tid < nn) {
        const float *q = cbase + (long long)s.typed_atom[s.carbon[cL + tid]] * 3;
        s_np[tid] = make_float4(q[0], q[1], q[2], 0.f);
        s_nw[tid] = q[0] > ABSENT_X ? 0.0 : w0c[cL + tid];
    }
    __syncwarp();
    double wf = 0.0;
    if (tid < nn) {
        double f = 0.0;
        const float4 me = s_np[tid];
        if (me.x <= ABSENT_X) {
            for (int t = 0; t < n_tiles; ++t) f += partial_new[((long long)c * n_tiles + t) * EXT_NEW_C + tid];
            for (int n = 0; n < nn; ++n) {
                if (n == tid || s_nw[n] == 0.0) continue;
                const float4 q = s_np[n];
                if (dist2_fast(me.x, me.y, me.z, q.x, q.y, q.z) < HPMF_CUTOFF * HPMF_CUTOFF)
                    f += s_nw[n] * gauss3(s, __fsqrt_rn(norm2_f32(me.x, me.y, me.z, q.x, q.y, q.z)));
            }
        }
        f0[(long long)c * s.n_carbon + cL + tid] = f;
        wf = s_nw[tid] * f;
    }
    // lane 0 adds everything up in a fixed order
    double sum = 0.0;
    for (int n = 0; n < nn; ++n) { const double x = __shfl_sync(0xffffffffu, wf, n); sum += x; }
    if (tid == 0) {
        for (int t = 0; t < n_tiles; ++t) sum += partial[(long long)c * n_tiles + t];
        v0[c] = 0.5 * sum;
    }
}

// positions of the environment's carbons, packed: one 16-byte load per carbon in the per-trial kernel
__global__ void hpmf_pack_carbons_kernel(HpmfDev s, const float *__restrict__ coords, int c1, float4 *__restrict__ cpos) {
    const int ck = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (ck >= c1) return;
    const float *p = coords + ((long long)c * s.n_atoms + s.typed_atom[s.carbon[ck]]) * 3;
    cpos[(long long)c * s.n_carbon + ck] = make_float4(p[0], p[1], p[2], 0.f);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#define HPMF_STATES 5          // [0]: the public mcsce_hpmf_trials API; [1..4]: the conformer groups (lanes) of mcsce_grow

// per-conformer environment state of one growth in flight (grow-only buffers)
struct EnvState {
    double *d_env_state = nullptr;          // sa_env [C*n_typed] | w0 [C*n_carbon] | f0 [C*n_carbon] | v0 [C]
    size_t env_state_cap = 0;
    double *d_partial = nullptr;
    size_t partial_cap = 0;
    void *d_ext = nullptr;                  // ext_pos | ext_del | ext_ck | ext_n
    size_t ext_cap = 0;
    double *d_partial_new = nullptr;        // [C*tiles*EXT_NEW_C] per-tile sums of the new carbons' fields
    float4 *d_cpos = nullptr;               // [C*n_carbon] packed carbon positions
    float4 *d_scr_pos = nullptr;            // [C*n_carbon] exposed carbons beyond the per-conformer kernel's shared-memory stage
    double *d_scr_w = nullptr;
    // carried state: the level (first atom not yet part of the state) the buffers describe
    int state_level = 0, state_nconf = 0;
    const float *state_coords = nullptr;
};

struct mcsce_hpmf_s {
    int device = 0;
    HpmfDev dev{};
    std::vector<void *> owned;
    double *d_partial = nullptr;            // whole-structure evaluations (mcsce_hpmf_energy)
    size_t partial_cap = 0;
    // per-trial evaluation: host copies of the index tables
    std::vector<int> h_typed_atom, h_carbon;
    const int *d_carbon_ord = nullptr;
    int *d_overflow = nullptr;
    bool carry = false;
    int growth_trial_kernel = 1;            // 1: per-conformer kernel inside mcsce_grow (hpmf_trials_conf_kernel); 0: per-trial kernel
    EnvState es[HPMF_STATES];
    // inside mcsce_grow: per-residue lists of typed environment atoms within SASA range of the residue's side chain
    int *d_near = nullptr, *d_near_c = nullptr;   // typed atoms in SASA range / carbon ordinals in Gaussian range
    std::vector<int> near_off, near_c_off;  // [n_res + 1] offsets into d_near / d_near_c
    long long near_version = -1;
    double *d_sasa1 = nullptr;              // scratch of the level-0 evaluation (one structure)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool timed = false;
};

template <typename T>
static int upload(mcsce_hpmf_s *h, const T *src, size_t n, const T **dst) {
    void *p = nullptr;
    HCK(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    h->owned.push_back(p);
    if (n) HCK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return 0;
}

extern "C" int mcsce_hpmf_create(mcsce_hpmf *out, int device, const mcsce_hpmf_desc *d) {
    if (!out || !d) return mcsce_fail_shared("mcsce_hpmf_create: null argument");
    if (d->n_types < 1 || d->n_types > MAX_SASA_TYPES) return mcsce_fail_shared("mcsce_hpmf_create: n_types out of range");
    if (d->n_typed < 0 || d->n_typed > d->n_atoms) return mcsce_fail_shared("mcsce_hpmf_create: n_typed out of range");
    HCK(cudaSetDevice(device));
    mcsce_hpmf_s *h = new mcsce_hpmf_s();
    h->device = device;
    const int nt = d->n_types;
    std::vector<double> thr((size_t)nt * nt), piRRs(nt), S(nt);
    const double pi = 3.141592653589793;
    double thr_max = 0.0;
    for (int a = 0; a < nt; ++a) {
        piRRs[a] = pi * (d->type_R[a] + d->r_solvent);                                   // np.pi * (Ri + Rs)
        const double rr = d->type_R[a] + d->r_solvent;
        S[a] = 4 * pi * (rr * rr);                                                        // 4 * np.pi * (Ri + Rs) ** 2
        for (int b = 0; b < nt; ++b) {
            thr[(size_t)a * nt + b] = d->type_R[a] + d->type_R[b] + 2 * d->r_solvent;     // Ri + Rj + 2 * Rs
            if (thr[(size_t)a * nt + b] > thr_max) thr_max = thr[(size_t)a * nt + b];
        }
    }
    std::vector<int> carbon;
    for (int i = 0; i < d->n_typed; ++i) {
        const int t = d->type_of[i];
        if (t < 0 || t >= nt) { delete h; return mcsce_fail_shared("mcsce_hpmf_create: type index out of range"); }
        if (d->typed_atom[i] < 0 || d->typed_atom[i] >= d->n_atoms) { delete h; return mcsce_fail_shared("mcsce_hpmf_create: typed_atom out of range"); }
        if (d->type_is_carbon[t]) carbon.push_back(i);
    }
    HpmfDev &v = h->dev;
    v.n_atoms = d->n_atoms; v.n_typed = d->n_typed; v.n_types = nt; v.n_carbon = (int)carbon.size();
    if (upload(h, d->typed_atom, (size_t)d->n_typed, &v.typed_atom)) return -1;
    if (upload(h, d->type_of, (size_t)d->n_typed, &v.type_of)) return -1;
    if (upload(h, d->bonded, (size_t)d->n_typed * MAXB, &v.bonded)) return -1;
    if (upload(h, carbon.data(), carbon.size(), &v.carbon)) return -1;
    {
        std::vector<int> ord((size_t)d->n_typed, -1);
        for (size_t k = 0; k < carbon.size(); ++k) ord[carbon[k]] = (int)k;
        if (upload(h, ord.data(), ord.size(), &h->d_carbon_ord)) return -1;
        h->h_typed_atom.assign(d->typed_atom, d->typed_atom + d->n_typed);
        h->h_carbon = carbon;
        for (int i = 1; i < d->n_typed; ++i)
            if (d->typed_atom[i] <= d->typed_atom[i - 1]) { delete h; return mcsce_fail_shared("mcsce_hpmf_create: typed_atom must be ascending"); }
        void *po = nullptr;
        HCK(cudaMalloc(&po, sizeof(int)));
        HCK(cudaMemset(po, 0, sizeof(int)));
        h->owned.push_back(po);
        h->d_overflow = (int *)po;
    }
    if (upload(h, thr.data(), thr.size(), &v.thr)) return -1;
    if (upload(h, d->type_R, (size_t)nt, &v.R)) return -1;
    if (upload(h, d->type_P, (size_t)nt, &v.P)) return -1;
    if (upload(h, piRRs.data(), (size_t)nt, &v.piRRs)) return -1;
    if (upload(h, S.data(), (size_t)nt, &v.S)) return -1;
    v.pij_bonded = d->pij_bonded; v.pij_non_bonded = d->pij_non_bonded;
    v.pre2 = (float)(thr_max * thr_max * 1.001);
    for (int k = 0; k < 3; ++k) {
        v.C[k] = d->hpmf_c[k]; v.W[k] = d->hpmf_w[k]; v.H[k] = d->hpmf_h[k];
        v.Cf[k] = (float)d->hpmf_c[k]; v.iWf[k] = (float)(1.0 / d->hpmf_w[k]); v.Hf[k] = (float)d->hpmf_h[k];
    }
    v.sa_threshold = d->sa_threshold;
    for (int k = 0; k < 3; ++k) HCK(cudaEventCreate(&h->ev[k]));
    *out = h;
    return 0;
}

extern "C" int mcsce_hpmf_destroy(mcsce_hpmf h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (void *p : h->owned) cudaFree(p);
    if (h->d_partial) cudaFree(h->d_partial);
    for (auto &e : h->es) {
        if (e.d_env_state) cudaFree(e.d_env_state);
        if (e.d_partial) cudaFree(e.d_partial);
        if (e.d_ext) cudaFree(e.d_ext);
        if (e.d_partial_new) cudaFree(e.d_partial_new);
        if (e.d_cpos) cudaFree(e.d_cpos);
        if (e.d_scr_pos) cudaFree(e.d_scr_pos);
        if (e.d_scr_w) cudaFree(e.d_scr_w);
    }
    if (h->d_near) cudaFree(h->d_near);
    if (h->d_near_c) cudaFree(h->d_near_c);
    if (h->d_sasa1) cudaFree(h->d_sasa1);
    for (int k = 0; k < 3; ++k) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    delete h;
    return 0;
}

static int launch_sasa(mcsce_hpmf h, int n_struct, const float *d_coords, double *d_sasa, cudaStream_t st, int nt_lim = -1) {
    if (nt_lim < 0) nt_lim = h->dev.n_typed;
    const HpmfDev &v = h->dev;
    if (v.n_typed == 0) return 0;
    const int tiles = (v.n_typed + HT - 1) / HT;
    for (int c0 = 0; c0 < n_struct; c0 += 65535) {
        const int nc = n_struct - c0 < 65535 ? n_struct - c0 : 65535;
        sasa_kernel<<<dim3(tiles, nc), HT, 0, st>>>(v, d_coords + (size_t)c0 * v.n_atoms * 3, d_sasa + (size_t)c0 * v.n_typed,
                                                    nt_lim);
    }
    HCK(cudaGetLastError());
    return 0;
}

extern "C" int mcsce_sasa(mcsce_hpmf h, int n_struct, const float *d_coords, double *d_sasa, void *stream) {
    if (!h || !d_coords || !d_sasa) return mcsce_fail_shared("mcsce_sasa: null argument");
    if (n_struct <= 0) return 0;
    HCK(cudaSetDevice(h->device));
    return launch_sasa(h, n_struct, d_coords, d_sasa, (cudaStream_t)stream);
}

extern "C" int mcsce_hpmf_energy(mcsce_hpmf h, int n_struct, const float *d_coords, double *d_sasa, double *d_v,
                                 void *stream) {
    if (!h || !d_coords || !d_sasa || !d_v) return mcsce_fail_shared("mcsce_hpmf_energy: null argument");
    if (n_struct <= 0) return 0;
    HCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const HpmfDev &v = h->dev;
    const int tiles = v.n_carbon > 0 ? (v.n_carbon + HT - 1) / HT : 1;
    const size_t need = (size_t)n_struct * tiles;
    if (need > h->partial_cap) {
        HCK(cudaStreamSynchronize(st));
        if (h->d_partial) HCK(cudaFree(h->d_partial));
        h->d_partial = nullptr; h->partial_cap = 0;
        HCK(cudaMalloc(&h->d_partial, need * sizeof(double)));
        h->partial_cap = need;
    }
    HCK(cudaEventRecord(h->ev[0], st));
    if (launch_sasa(h, n_struct, d_coords, d_sasa, st)) return -1;
    HCK(cudaEventRecord(h->ev[1], st));
    for (int c0 = 0; c0 < n_struct; c0 += 65535) {
        const int nc = n_struct - c0 < 65535 ? n_struct - c0 : 65535;
        hpmf_pair_kernel<<<dim3(tiles, nc), HT, 0, st>>>(v, d_coords + (size_t)c0 * v.n_atoms * 3,
                                                         d_sasa + (size_t)c0 * v.n_typed, h->d_partial + (size_t)c0 * tiles,
                                                         v.n_carbon, nullptr, nullptr);
    }
    HCK(cudaGetLastError());
    HCK(cudaEventRecord(h->ev[2], st));
    hpmf_finish_kernel<<<(n_struct + 127) / 128, 128, 0, st>>>(n_struct, tiles, h->d_partial, d_v);
    HCK(cudaGetLastError());
    h->timed = true;
    return 0;
}


// Per-trial V_hpmf of one growth step.  `es`: the environment state this growth carries; `carry`: extend it from the
// previous step instead of re-deriving it.  Buffers grow here (stream synchronised when they do: only at a growth's first step).
struct TrialStepExtra {
    float ca[3] = {0.f, 0.f, 0.f}; float reach = -1.f;                  // the residue's CA and reach bound (< 0: lists are not filtered)
    const float4 *trial4 = nullptr; long long trial4_stride = 0;      // the growth's own trial buffer instead of d_trial
    const int *near = nullptr; int n_near = 0;                          // typed atoms in SASA range of the residue (complete list)
    const int *clist = nullptr; int n_clist = 0;                        // carbons in Gaussian range of the residue (complete list)
    int *overflow_conf = nullptr;                                       // [n_conf] set where a capacity was exceeded
};

static int hpmf_trials_impl(mcsce_hpmf h, EnvState &es, bool carry, int n_conf, const float *d_coords, int first_new, int n_new,
                            int n_trials, const float *d_trial, const uint8_t *d_mask, double *d_v, cudaStream_t st,
                            const TrialStepExtra &x) {
    const HpmfDev &v = h->dev;
    const std::vector<int> &ta = h->h_typed_atom;
    const int e1 = (int)(std::lower_bound(ta.begin(), ta.end(), first_new) - ta.begin());
    const int e2 = (int)(std::lower_bound(ta.begin(), ta.end(), first_new + n_new) - ta.begin());
    const int c1 = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), e1) - h->h_carbon.begin());
    const int c2 = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), e2) - h->h_carbon.begin());
    const int m = e2 - e1, mc = c2 - c1;
    if (m > MAX_NEW_T || mc > MAX_NEW_C) return mcsce_fail_shared("mcsce_hpmf_trials: too many typed atoms in the new residue");
    const int tiles = c1 > 0 ? (c1 + HT - 1) / HT : 1;
    const size_t n_state = (size_t)n_conf * ((size_t)v.n_typed + 2 * (size_t)v.n_carbon + 1);
    const size_t need_partial = (size_t)n_conf * ((v.n_carbon + HT - 1) / HT + 1);
    if (n_state > es.env_state_cap || need_partial > es.partial_cap) {
        HCK(cudaStreamSynchronize(st));
        if (n_state > es.env_state_cap) {
            es.state_level = 0;
            if (es.d_env_state) HCK(cudaFree(es.d_env_state));
            es.d_env_state = nullptr; es.env_state_cap = 0;
            HCK(cudaMalloc(&es.d_env_state, n_state * sizeof(double)));
            es.env_state_cap = n_state;
        }
        if (need_partial > es.partial_cap) {
            if (es.d_partial) HCK(cudaFree(es.d_partial));
            if (es.d_partial_new) HCK(cudaFree(es.d_partial_new));
            es.d_partial = nullptr; es.d_partial_new = nullptr; es.partial_cap = 0;
            HCK(cudaMalloc(&es.d_partial, need_partial * sizeof(double)));
            HCK(cudaMalloc(&es.d_partial_new, need_partial * EXT_NEW_C * sizeof(double)));
            es.partial_cap = need_partial;
        }
        if (es.d_cpos) HCK(cudaFree(es.d_cpos));
        if (es.d_scr_pos) HCK(cudaFree(es.d_scr_pos));
        if (es.d_scr_w) HCK(cudaFree(es.d_scr_w));
        es.d_cpos = nullptr; es.d_scr_pos = nullptr; es.d_scr_w = nullptr;
        HCK(cudaMalloc(&es.d_cpos, (size_t)n_conf * std::max(1, v.n_carbon) * sizeof(float4)));
        HCK(cudaMalloc(&es.d_scr_pos, (size_t)n_conf * std::max(1, v.n_carbon) * sizeof(float4)));
        HCK(cudaMalloc(&es.d_scr_w, (size_t)n_conf * std::max(1, v.n_carbon) * sizeof(double)));
    }
    double *sa_env = es.d_env_state;
    double *w0 = sa_env + (size_t)n_conf * v.n_typed;
    double *f0 = w0 + (size_t)n_conf * v.n_carbon;
    double *v0 = f0 + (size_t)n_conf * v.n_carbon;
    // the environment alone (atoms before first_new), once per conformer: extended from the previous step's state
    // when the caller allowed it (mcsce_hpmf_carry) and the call continues that growth, else derived from scratch
    bool extended = false;
    if (carry && es.state_level > 0 && es.state_nconf == n_conf && es.state_coords == d_coords &&
        es.state_level <= first_new) {
        const int eL = (int)(std::lower_bound(ta.begin(), ta.end(), es.state_level) - ta.begin());
        const int cL = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), eL) - h->h_carbon.begin());
        const int mL = e1 - eL;
        if (mL == 0) extended = true;                           // only untyped atoms were added: nothing changes
        else if (mL <= MAX_NEW_T) {
            const size_t per = (size_t)EXT_LIST * (sizeof(float4) + sizeof(double) + sizeof(int)) + sizeof(int);
            if ((size_t)n_conf * per > es.ext_cap) {
                HCK(cudaStreamSynchronize(st));
                if (es.d_ext) HCK(cudaFree(es.d_ext));
                es.d_ext = nullptr; es.ext_cap = 0;
                HCK(cudaMalloc(&es.d_ext, (size_t)n_conf * per));
                es.ext_cap = (size_t)n_conf * per;
            }
            float4 *ext_pos = (float4 *)es.d_ext;
            double *ext_del = (double *)(ext_pos + (size_t)n_conf * EXT_LIST);
            int *ext_ck = (int *)(ext_del + (size_t)n_conf * EXT_LIST);
            int *ext_n = ext_ck + (size_t)n_conf * EXT_LIST;
            hpmf_extend_kernel<<<n_conf, HT, 0, st>>>(v, d_coords, eL, mL, h->d_carbon_ord, sa_env, w0, ext_pos, ext_del, ext_ck,
                                                      ext_n, h->d_overflow, x.overflow_conf);
            const int tilesL = cL > 0 ? (cL + HT - 1) / HT : 1;
            hpmf_field_update_kernel<<<dim3(tilesL, n_conf), HT, 0, st>>>(v, d_coords, cL, c1, w0, f0, ext_pos, ext_del, ext_ck,
                                                                          ext_n, es.d_partial, es.d_partial_new);
            hpmf_field_finish_kernel<<<n_conf, 32, 0, st>>>(v, d_coords, cL, c1, tilesL, w0, f0, es.d_partial, es.d_partial_new, v0);
            extended = true;
        }
    }
    if (!extended) {
        if (launch_sasa(h, n_conf, d_coords, sa_env, st, e1)) return -1;
        hpmf_pair_kernel<<<dim3(tiles, n_conf), HT, 0, st>>>(v, d_coords, sa_env, es.d_partial, c1, w0, f0);
        hpmf_finish_kernel<<<(n_conf + 127) / 128, 128, 0, st>>>(n_conf, tiles, es.d_partial, v0);
    }
    es.state_level = first_new; es.state_nconf = n_conf; es.state_coords = d_coords;
    if (c1 > 0) hpmf_pack_carbons_kernel<<<dim3((c1 + 127) / 128, n_conf), 128, 0, st>>>(v, d_coords, c1, es.d_cpos);
    // every trial on top of it
    if (x.trial4 && h->growth_trial_kernel && n_trials <= CT_TRIALS) {
        ConfTrialArgs a{};
        a.s = v; a.coords = d_coords; a.trial4 = x.trial4; a.trial4_stride = x.trial4_stride; a.mask = d_mask;
        a.first_new = first_new; a.n_new = n_new; a.n_trials = n_trials; a.e1 = e1; a.m = m; a.c1 = c1; a.mc = mc;
        a.carbon_ord = h->d_carbon_ord; a.sa_env = sa_env; a.w0 = w0; a.f0 = f0; a.v0 = v0; a.out = d_v;
        a.overflow = h->d_overflow; a.overflow_conf = x.overflow_conf;
        a.near = x.near; a.n_near = x.n_near; a.clist = x.clist; a.n_clist = x.n_clist;
        a.cpos = es.d_cpos; a.scratch_pos = es.d_scr_pos; a.scratch_w = es.d_scr_w;
        a.cax = x.ca[0]; a.cay = x.ca[1]; a.caz = x.ca[2]; a.r2_near = -1.f; a.r2_carb = -1.f;
        if (x.reach >= 0.f) {
            const float rn = x.reach + sqrtf(v.pre2) + 0.01f, rc = x.reach + HPMF_CUTOFF + 0.01f;
            a.r2_near = rn * rn; a.r2_carb = rc * rc;
        }
        static bool attr_set = false;
        if (!attr_set) {
            HCK(cudaFuncSetAttribute(hpmf_trials_conf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConfTrialSmem)));
            attr_set = true;
        }
        hpmf_trials_conf_kernel<<<n_conf, CT_THREADS, sizeof(ConfTrialSmem), st>>>(a);
    } else {
        static int minb = -1;                 // tuning knob: resident blocks per SM the kernel is compiled for
        if (minb < 0) { const char *e = getenv("MCSCE_B200_HPMF_MINB"); minb = e ? atoi(e) : TRIAL_MINB_DEFAULT; }
        const dim3 grid(n_trials, n_conf);
#define LAUNCH_TRIAL(B) hpmf_trial_kernel<B><<<grid, HT, 0, st>>>(v, d_coords, d_trial, d_mask, first_new, n_new, n_trials, \
            e1, m, c1, mc, h->d_carbon_ord, sa_env, w0, f0, v0, d_v, h->d_overflow, x.trial4, x.trial4_stride, x.near, x.n_near, \
            x.overflow_conf, es.d_cpos, x.clist, x.n_clist)
        if (minb >= 8) LAUNCH_TRIAL(8); else if (minb >= 6) LAUNCH_TRIAL(6); else LAUNCH_TRIAL(4);
#undef LAUNCH_TRIAL
    }
    HCK(cudaGetLastError());
    return 0;
}

extern "C" int mcsce_hpmf_trials(mcsce_hpmf h, int n_conf, const float *d_coords, int first_new, int n_new, int n_trials,
                                 const float *d_trial, const uint8_t *d_mask, double *d_v, void *stream) {
    if (!h || !d_coords || !d_trial || !d_v) return mcsce_fail_shared("mcsce_hpmf_trials: null argument");
    if (n_conf <= 0 || n_trials <= 0) return 0;
    if (n_conf > 65535) return mcsce_fail_shared("mcsce_hpmf_trials: at most 65535 conformers per call");
    const HpmfDev &v = h->dev;
    if (first_new < 0 || n_new <= 0 || first_new + n_new > v.n_atoms) return mcsce_fail_shared("mcsce_hpmf_trials: new-atom range out of bounds");
    HCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (hpmf_trials_impl(h, h->es[0], h->carry, n_conf, d_coords, first_new, n_new, n_trials, d_trial, d_mask, d_v, st,
                         TrialStepExtra{})) return -1;
    int over = 0;
    HCK(cudaMemcpyAsync(&over, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    HCK(cudaStreamSynchronize(st));
    if (over) {
        HCK(cudaMemset(h->d_overflow, 0, sizeof(int)));
        h->es[0].state_level = 0;
        return mcsce_fail_shared("mcsce_hpmf_trials: more than 512 environment carbons change weight under one trial");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// the term inside mcsce_grow (called from mcsce_kernels.cu; not part of the C ABI)
// ------------------------------------------------------------------------------------------------
// Lists of typed environment atoms per grown residue, from growth-order atom lists (mcsce_kernels.cu builds those from
// rigorous reach bounds): untyped atoms (carbon-bound hydrogens) are dropped, indices become typed indices.
int mcsce_hpmf_set_near_lists(mcsce_hpmf h, long long version, int n_res, const int *off, const int *atoms, const int *off2,
                              const int *atoms2) {
    if (!h) return mcsce_fail_shared("null hpmf handle");
    HCK(cudaSetDevice(h->device));
    const std::vector<int> &ta = h->h_typed_atom;
    std::vector<int> ord((size_t)h->dev.n_typed, -1);
    for (size_t k = 0; k < h->h_carbon.size(); ++k) ord[h->h_carbon[k]] = (int)k;
    auto convert = [&](const int *o, const int *at, bool carbons, std::vector<int> &offs, std::vector<int> &out) {
        out.clear(); out.reserve(o[n_res]);
        offs.assign(n_res + 1, 0);
        for (int i = 0; i < n_res; ++i) {
            offs[i] = (int)out.size();
            for (int q = o[i]; q < o[i + 1]; ++q) {
                auto it = std::lower_bound(ta.begin(), ta.end(), at[q]);
                if (it == ta.end() || *it != at[q]) continue;
                const int k = (int)(it - ta.begin());
                if (!carbons) out.push_back(k);
                else if (ord[k] >= 0) out.push_back(ord[k]);
            }
            std::sort(out.begin() + offs[i], out.end());
        }
        offs[n_res] = (int)out.size();
    };
    std::vector<int> l1, l2;
    convert(off, atoms, false, h->near_off, l1);
    convert(off2, atoms2, true, h->near_c_off, l2);
    HCK(cudaDeviceSynchronize());
    if (h->d_near) HCK(cudaFree(h->d_near));
    if (h->d_near_c) HCK(cudaFree(h->d_near_c));
    h->d_near = nullptr; h->d_near_c = nullptr;
    HCK(cudaMalloc(&h->d_near, sizeof(int) * std::max<size_t>(1, l1.size())));
    HCK(cudaMalloc(&h->d_near_c, sizeof(int) * std::max<size_t>(1, l2.size())));
    if (!l1.empty()) HCK(cudaMemcpy(h->d_near, l1.data(), sizeof(int) * l1.size(), cudaMemcpyHostToDevice));
    if (!l2.empty()) HCK(cudaMemcpy(h->d_near_c, l2.data(), sizeof(int) * l2.size(), cudaMemcpyHostToDevice));
    h->near_version = version;
    return 0;
}

long long mcsce_hpmf_near_version(mcsce_hpmf h) { return h ? h->near_version : -1; }
// largest distance at which two typed atoms shadow each other: max (R_i + R_j + 2 R_s), as padded for the float32 prefilter
double mcsce_hpmf_sasa_range(mcsce_hpmf h) { return h ? sqrt((double)h->dev.pre2) : 0.0; }
double mcsce_hpmf_pair_range(mcsce_hpmf) { return (double)HPMF_CUTOFF; }
int mcsce_hpmf_n_atoms(mcsce_hpmf h) { return h ? h->dev.n_atoms : -1; }

// Start of a growth on lane `lane`: forget the carried state; d_v_init[0] = V_hpmf of the first n_input atoms alone (the level-0
// energy every conformer starts from, side_chain_builder.py:149), evaluated on one structure of d_coords.
int mcsce_hpmf_growth_begin(mcsce_hpmf h, int lane, const float *d_coords, int n_input, double *d_v_init, cudaStream_t st) {
    if (!h || lane < 0 || lane + 1 >= HPMF_STATES) return mcsce_fail_shared("mcsce_hpmf_growth_begin: bad arguments");
    HCK(cudaSetDevice(h->device));
    h->es[lane + 1].state_level = 0;
    if (!d_v_init) return 0;
    const HpmfDev &v = h->dev;
    const std::vector<int> &ta = h->h_typed_atom;
    const int e0 = (int)(std::lower_bound(ta.begin(), ta.end(), n_input) - ta.begin());
    const int c0 = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), e0) - h->h_carbon.begin());
    const int tiles = c0 > 0 ? (c0 + HT - 1) / HT : 1;
    if (!h->d_sasa1) HCK(cudaMalloc(&h->d_sasa1, sizeof(double) * (std::max(1, v.n_typed) + (v.n_carbon + HT - 1) / HT + 2)));
    double *partial = h->d_sasa1 + std::max(1, v.n_typed);
    if (launch_sasa(h, 1, d_coords, h->d_sasa1, st, e0)) return -1;
    hpmf_pair_kernel<<<dim3(tiles, 1), HT, 0, st>>>(v, d_coords, h->d_sasa1, partial, c0, nullptr, nullptr);
    hpmf_finish_kernel<<<1, 128, 0, st>>>(1, tiles, partial, d_v_init);
    HCK(cudaGetLastError());
    return 0;
}

// One residue of a growth: V_hpmf(environment + trial) for the masked (conformer, trial) pairs, into d_v [n_conf*n_trials].
int mcsce_hpmf_growth_step(mcsce_hpmf h, int lane, int residue, int n_conf, const float *d_coords, int first_new, int n_new,
                           int n_trials, const float4 *d_trial4, long long trial4_stride, const uint8_t *d_mask, double *d_v,
                           int *d_overflow_conf, cudaStream_t st, const float *ca, float reach) {
    if (!h || lane < 0 || lane + 1 >= HPMF_STATES) return mcsce_fail_shared("mcsce_hpmf_growth_step: bad arguments");
    HCK(cudaSetDevice(h->device));
    TrialStepExtra x;
    x.trial4 = d_trial4; x.trial4_stride = trial4_stride; x.overflow_conf = d_overflow_conf;
    if (ca) { x.ca[0] = ca[0]; x.ca[1] = ca[1]; x.ca[2] = ca[2]; x.reach = reach; }
    static int no_lists = -1;                 // diagnostic knob: visit every environment atom instead of the residue's lists
    if (no_lists < 0) { const char *e = getenv("MCSCE_B200_HPMF_NO_LISTS"); no_lists = e ? atoi(e) : 0; }
    if (h->d_near && !no_lists && residue + 1 < (int)h->near_off.size()) {
        x.near = h->d_near + h->near_off[residue]; x.n_near = h->near_off[residue + 1] - h->near_off[residue];
        x.clist = h->d_near_c + h->near_c_off[residue]; x.n_clist = h->near_c_off[residue + 1] - h->near_c_off[residue];
    }
    return hpmf_trials_impl(h, h->es[lane + 1], true, n_conf, d_coords, first_new, n_new, n_trials, nullptr, d_mask, d_v, st, x);
}

extern "C" int mcsce_hpmf_set_option(mcsce_hpmf h, int key, int value) {
    if (!h) return mcsce_fail_shared("mcsce_hpmf_set_option: null handle");
    if (key == 0) { h->growth_trial_kernel = value ? 1 : 0; return 0; }
    return mcsce_fail_shared("mcsce_hpmf_set_option: unknown key");
}

extern "C" int mcsce_hpmf_carry(mcsce_hpmf h, int enable) {
    if (!h) return mcsce_fail_shared("mcsce_hpmf_carry: null handle");
    h->carry = enable != 0;
    h->es[0].state_level = 0;               // the next mcsce_hpmf_trials call derives the environment from scratch
    return 0;
}

extern "C" int mcsce_hpmf_last_ms(mcsce_hpmf h, double *ms) {
    if (!h || !ms) return mcsce_fail_shared("mcsce_hpmf_last_ms: null argument");
    if (!h->timed) return mcsce_fail_shared("mcsce_hpmf_last_ms: no mcsce_hpmf_energy call yet");
    HCK(cudaSetDevice(h->device));
    HCK(cudaEventSynchronize(h->ev[2]));
    float a = 0.f, b = 0.f;
    HCK(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
    HCK(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
    ms[0] = a; ms[1] = b;
    return 0;
}
