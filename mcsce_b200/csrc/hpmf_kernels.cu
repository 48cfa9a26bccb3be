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
// squared distance; Gaussian terms beyond 18 A (< 2e-22 each) are dropped.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
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
#define HPMF_CUTOFF 18.0f
#define ABSENT_X 1.0e14f       // x beyond this = atom not placed (the growth parks unplaced atoms at 1e15)

struct HpmfDev {
    int n_atoms, n_typed, n_types, n_carbon;
    const int *typed_atom, *type_of, *bonded, *carbon;   // carbon: [n_carbon] typed index
    const double *thr;                                   // [n_types*n_types] R_i + R_j + 2 R_s
    const double *R, *P, *piRRs, *S;                     // per type: R, P, pi (R + R_s), 4 pi (R + R_s)^2
    double pij_bonded, pij_non_bonded;
    float pre2;                                          // float32 prefilter on d^2: (largest threshold)^2, padded
    double C[3], W[3], H[3], sa_threshold;
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

// one factor of the product: atom i (P_i, S_i, R_i) shadowed by partner j (pi (R_j + R_s), R_j) at distance d < thr
__device__ __forceinline__ double sasa_term(double Pi, double Si, double Ri, double piRRs_j, double Rj, double thr, float d,
                                            double pij) {
    // dij + 1e-8 stays float32 in the reference (float32 array + Python scalar)
    const double dd = (double)__fadd_rn(d, 1e-8f);
    const double bij = __dmul_rn(__dmul_rn(piRRs_j, __dsub_rn(thr, (double)d)),
                                 __dadd_rn(1.0, __ddiv_rn(__dsub_rn(Rj, Ri), dd)));
    return __dsub_rn(1.0, __ddiv_rn(__dmul_rn(__dmul_rn(Pi, pij), bij), Si));
}

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

__device__ __forceinline__ double gauss3(const HpmfDev &s, double r) {
    const double t0 = (r - s.C[0]) / s.W[0], t1 = (r - s.C[1]) / s.W[1], t2 = (r - s.C[2]) / s.W[2];
    return s.H[0] * exp(-(t0 * t0)) + s.H[1] * exp(-(t1 * t1)) + s.H[2] * exp(-(t2 * t2));
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
        wi = (a > s.sa_threshold && xi <= ABSENT_X) ? tanh(a) : 0.0;
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
            const double w = (a > s.sa_threshold && p[0] <= ABSENT_X) ? tanh(a) : 0.0;
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
                const double r = (double)__fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z));
                acc += wj * gauss3(s, r);
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
        const double *__restrict__ v0, double *__restrict__ out, int *__restrict__ overflow) {
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
        const float *p = trial + (ct * n_new + (s.typed_atom[g] - first_new)) * 3;
        ta[tid] = make_float4(p[0], p[1], p[2], __int_as_float(s.type_of[g]));
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
    for (int k = tid; k < e1; k += HT) {
        const float *p = cbase + (long long)s.typed_atom[k] * 3;
        const float xk = p[0], yk = p[1], zk = p[2];
        if (xk > ABSENT_X) continue;
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
                const double w = a > s.sa_threshold ? tanh(a) : 0.0;
                const double del = w - w0c[ck];
                if (del != 0.0) {
                    acc += del * f0c[ck];
                    const int idx = atomicAdd(&ln, 1);
                    if (idx < MAX_LIST) { lpos[idx] = make_float4(xk, yk, zk, 0.f); ldel[idx] = del; }
                }
            }
        }
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
            tw[ck - c1] = a > s.sa_threshold ? tanh(a) : 0.0;
            tc[ck - c1] = make_float4(me.x, me.y, me.z, 0.f);
        }
    }
    __syncthreads();
    const int L = min(ln, MAX_LIST);
    if (tid == 0 && ln > MAX_LIST) atomicExch(overflow, 1);
    const float cut2 = HPMF_CUTOFF * HPMF_CUTOFF;
    // phase 3a: trial carbons x environment carbons (weights of the environment alone)
    for (int ck = tid; ck < c1; ck += HT) {
        const double wk = w0c[ck];
        if (wk == 0.0) continue;
        const float *p = cbase + (long long)s.typed_atom[s.carbon[ck]] * 3;
        const float xk = p[0], yk = p[1], zk = p[2];
        for (int n = 0; n < mc; ++n) {
            const double wn = tw[n];
            const float4 q = tc[n];
            if (wn != 0.0 && dist2_fast(xk, yk, zk, q.x, q.y, q.z) < cut2)
                acc += wn * wk * gauss3(s, (double)__fsqrt_rn(norm2_f32(xk, yk, zk, q.x, q.y, q.z)));
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
                acc += wn * da * gauss3(s, (double)__fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, q.x, q.y, q.z)));
        }
        for (int b = a + 1; b < L; ++b) {
            const float4 pb = lpos[b];
            if (dist2_fast(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < cut2)
                acc += da * ldel[b] * gauss3(s, (double)__fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z)));
        }
    }
    // phase 3c: trial carbons among themselves
    if (tid < mc) {
        const float4 pa = tc[tid];
        for (int b = tid + 1; b < mc; ++b) {
            const float4 pb = tc[b];
            if (tw[tid] != 0.0 && tw[b] != 0.0 && dist2_fast(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < cut2)
                acc += tw[tid] * tw[b] * gauss3(s, (double)__fsqrt_rn(norm2_f32(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z)));
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
        int *__restrict__ ext_ck, int *__restrict__ ext_n, int *__restrict__ overflow) {
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
    double tp[MAX_NEW_T];
#pragma unroll
    for (int n = 0; n < MAX_NEW_T; ++n) tp[n] = 1.0;
    for (int k = tid; k < eL; k += HT) {
        const float *p = cbase + (long long)s.typed_atom[k] * 3;
        const float xk = p[0], yk = p[1], zk = p[2];
        if (xk > ABSENT_X) continue;
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
            const int ck = carbon_ord[k];
            if (ck >= 0) {
                const double w = a > s.sa_threshold ? tanh(a) : 0.0;
                const double del = w - w0c[ck];
                if (del != 0.0) {
                    w0c[ck] = w;
                    const int idx = atomicAdd(&ln, 1);
                    if (idx < EXT_LIST) {
                        const long long o = (long long)c * EXT_LIST + idx;
                        ext_pos[o] = make_float4(xk, yk, zk, 0.f); ext_del[o] = del; ext_ck[o] = ck;
                    }
                }
            }
        }
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
        if (ck >= 0) w0c[ck] = a > s.sa_threshold ? tanh(a) : 0.0;
    }
    if (tid == 0) {
        ext_n[c] = min(ln, EXT_LIST);
        if (ln > EXT_LIST) atomicExch(overflow, 1);
    }
}

// fields and V of the extended environment: carbons [0, cL) gain the changed weights' and the new carbons' terms,
// the new carbons [cL, cF) get their full sums; partial[conformer][tile] = sum_i w_i F_i of the tile
__global__ void __launch_bounds__(HT) hpmf_field_update_kernel(
        HpmfDev s, const float *__restrict__ coords, int cL, int cF, const double *__restrict__ w0, double *__restrict__ f0,
        const float4 *__restrict__ ext_pos, const double *__restrict__ ext_del, const int *__restrict__ ext_ck,
        const int *__restrict__ ext_n, double *__restrict__ partial) {
    __shared__ double red[HT];
    const int c = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * HT + tid;
    const float *cbase = coords + (long long)c * s.n_atoms * 3;
    const double *w0c = w0 + (long long)c * s.n_carbon;
    double *f0c = f0 + (long long)c * s.n_carbon;
    const float cut2 = HPMF_CUTOFF * HPMF_CUTOFF;
    double contrib = 0.0;
    if (i < cF) {
        const float *p = cbase + (long long)s.typed_atom[s.carbon[i]] * 3;
        const float xi = p[0], yi = p[1], zi = p[2];
        double f = 0.0;
        if (xi <= ABSENT_X) {
            if (i < cL) {
                f = f0c[i];
                const int L = ext_n[c];
                const long long o = (long long)c * EXT_LIST;
                for (int l = 0; l < L; ++l) {
                    if (ext_ck[o + l] == i) continue;
                    const float4 q = ext_pos[o + l];
                    if (dist2_fast(xi, yi, zi, q.x, q.y, q.z) < cut2)
                        f += ext_del[o + l] * gauss3(s, (double)__fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z)));
                }
                for (int n = cL; n < cF; ++n) {
                    const double wn = w0c[n];
                    if (wn == 0.0) continue;
                    const float *q = cbase + (long long)s.typed_atom[s.carbon[n]] * 3;
                    if (dist2_fast(xi, yi, zi, q[0], q[1], q[2]) < cut2)
                        f += wn * gauss3(s, (double)__fsqrt_rn(norm2_f32(xi, yi, zi, q[0], q[1], q[2])));
                }
            } else {
                for (int j = 0; j < cF; ++j) {
                    const double wj = w0c[j];
                    if (wj == 0.0 || j == i) continue;
                    const float *q = cbase + (long long)s.typed_atom[s.carbon[j]] * 3;
                    if (dist2_fast(xi, yi, zi, q[0], q[1], q[2]) < cut2)
                        f += wj * gauss3(s, (double)__fsqrt_rn(norm2_f32(xi, yi, zi, q[0], q[1], q[2])));
                }
            }
        }
        f0c[i] = f;
        contrib = w0c[i] * f;
    }
    red[tid] = contrib;
    __syncthreads();
    for (int k = HT / 2; k > 0; k >>= 1) {
        if (tid < k) red[tid] += red[tid + k];
        __syncthreads();
    }
    if (tid == 0) partial[(long long)c * gridDim.x + blockIdx.x] = red[0];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct mcsce_hpmf_s {
    int device = 0;
    HpmfDev dev{};
    std::vector<void *> owned;
    double *d_partial = nullptr;
    size_t partial_cap = 0;
    // per-trial evaluation: host copies of the index tables, per-conformer environment state (grow-only)
    std::vector<int> h_typed_atom, h_carbon;
    const int *d_carbon_ord = nullptr;
    double *d_env_state = nullptr;          // sa_env [C*n_typed] | w0 [C*n_carbon] | f0 [C*n_carbon] | v0 [C]
    size_t env_state_cap = 0;
    int *d_overflow = nullptr;
    // carried state (mcsce_hpmf_carry): the level (first atom not yet part of the state) the buffers describe
    bool carry = false;
    int state_level = 0, state_nconf = 0;
    const float *state_coords = nullptr;
    void *d_ext = nullptr;                  // ext_pos | ext_del | ext_ck | ext_n
    size_t ext_cap = 0;
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
    for (int k = 0; k < 3; ++k) { v.C[k] = d->hpmf_c[k]; v.W[k] = d->hpmf_w[k]; v.H[k] = d->hpmf_h[k]; }
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
    if (h->d_env_state) cudaFree(h->d_env_state);
    if (h->d_ext) cudaFree(h->d_ext);
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


extern "C" int mcsce_hpmf_trials(mcsce_hpmf h, int n_conf, const float *d_coords, int first_new, int n_new, int n_trials,
                                 const float *d_trial, const uint8_t *d_mask, double *d_v, void *stream) {
    if (!h || !d_coords || !d_trial || !d_v) return mcsce_fail_shared("mcsce_hpmf_trials: null argument");
    if (n_conf <= 0 || n_trials <= 0) return 0;
    if (n_conf > 65535) return mcsce_fail_shared("mcsce_hpmf_trials: at most 65535 conformers per call");
    const HpmfDev &v = h->dev;
    if (first_new < 0 || n_new <= 0 || first_new + n_new > v.n_atoms) return mcsce_fail_shared("mcsce_hpmf_trials: new-atom range out of bounds");
    HCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const std::vector<int> &ta = h->h_typed_atom;
    const int e1 = (int)(std::lower_bound(ta.begin(), ta.end(), first_new) - ta.begin());
    const int e2 = (int)(std::lower_bound(ta.begin(), ta.end(), first_new + n_new) - ta.begin());
    const int c1 = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), e1) - h->h_carbon.begin());
    const int c2 = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), e2) - h->h_carbon.begin());
    const int m = e2 - e1, mc = c2 - c1;
    if (m > MAX_NEW_T || mc > MAX_NEW_C) return mcsce_fail_shared("mcsce_hpmf_trials: too many typed atoms in the new residue");
    const int tiles = c1 > 0 ? (c1 + HT - 1) / HT : 1;
    const size_t n_state = (size_t)n_conf * ((size_t)v.n_typed + 2 * (size_t)v.n_carbon + 1);
    if (n_state > h->env_state_cap || (size_t)n_conf * tiles > h->partial_cap) {
        HCK(cudaStreamSynchronize(st));
        if (n_state > h->env_state_cap) {
            h->state_level = 0;
            if (h->d_env_state) HCK(cudaFree(h->d_env_state));
            h->d_env_state = nullptr; h->env_state_cap = 0;
            HCK(cudaMalloc(&h->d_env_state, n_state * sizeof(double)));
            h->env_state_cap = n_state;
        }
        const size_t need = (size_t)n_conf * ((v.n_carbon + HT - 1) / HT + 1);
        if (need > h->partial_cap) {
            if (h->d_partial) HCK(cudaFree(h->d_partial));
            h->d_partial = nullptr; h->partial_cap = 0;
            HCK(cudaMalloc(&h->d_partial, need * sizeof(double)));
            h->partial_cap = need;
        }
    }
    double *sa_env = h->d_env_state;
    double *w0 = sa_env + (size_t)n_conf * v.n_typed;
    double *f0 = w0 + (size_t)n_conf * v.n_carbon;
    double *v0 = f0 + (size_t)n_conf * v.n_carbon;
    // the environment alone (atoms before first_new), once per conformer: extended from the previous step's state
    // when the caller allowed it (mcsce_hpmf_carry) and the call continues that growth, else derived from scratch
    bool extended = false;
    if (h->carry && h->state_level > 0 && h->state_nconf == n_conf && h->state_coords == d_coords &&
        h->state_level <= first_new) {
        const int eL = (int)(std::lower_bound(ta.begin(), ta.end(), h->state_level) - ta.begin());
        const int cL = (int)(std::lower_bound(h->h_carbon.begin(), h->h_carbon.end(), eL) - h->h_carbon.begin());
        const int mL = e1 - eL;
        if (mL == 0) extended = true;                           // only untyped atoms were added: nothing changes
        else if (mL <= MAX_NEW_T) {
            const size_t per = (size_t)EXT_LIST * (sizeof(float4) + sizeof(double) + sizeof(int)) + sizeof(int);
            if ((size_t)n_conf * per > h->ext_cap) {
                HCK(cudaStreamSynchronize(st));
                if (h->d_ext) HCK(cudaFree(h->d_ext));
                h->d_ext = nullptr; h->ext_cap = 0;
                HCK(cudaMalloc(&h->d_ext, (size_t)n_conf * per));
                h->ext_cap = (size_t)n_conf * per;
            }
            float4 *ext_pos = (float4 *)h->d_ext;
            double *ext_del = (double *)(ext_pos + (size_t)n_conf * EXT_LIST);
            int *ext_ck = (int *)(ext_del + (size_t)n_conf * EXT_LIST);
            int *ext_n = ext_ck + (size_t)n_conf * EXT_LIST;
            hpmf_extend_kernel<<<n_conf, HT, 0, st>>>(v, d_coords, eL, mL, h->d_carbon_ord, sa_env, w0, ext_pos, ext_del, ext_ck,
                                                      ext_n, h->d_overflow);
            hpmf_field_update_kernel<<<dim3(tiles, n_conf), HT, 0, st>>>(v, d_coords, cL, c1, w0, f0, ext_pos, ext_del, ext_ck,
                                                                         ext_n, h->d_partial);
            hpmf_finish_kernel<<<(n_conf + 127) / 128, 128, 0, st>>>(n_conf, tiles, h->d_partial, v0);
            extended = true;
        }
    }
    if (!extended) {
        if (launch_sasa(h, n_conf, d_coords, sa_env, st, e1)) return -1;
        hpmf_pair_kernel<<<dim3(tiles, n_conf), HT, 0, st>>>(v, d_coords, sa_env, h->d_partial, c1, w0, f0);
        hpmf_finish_kernel<<<(n_conf + 127) / 128, 128, 0, st>>>(n_conf, tiles, h->d_partial, v0);
    }
    h->state_level = first_new; h->state_nconf = n_conf; h->state_coords = d_coords;
    // every trial on top of it
    {
        static int minb = -1;                 // tuning knob: resident blocks per SM the kernel is compiled for
        if (minb < 0) { const char *e = getenv("MCSCE_B200_HPMF_MINB"); minb = e ? atoi(e) : TRIAL_MINB_DEFAULT; }
        const dim3 grid(n_trials, n_conf);
#define LAUNCH_TRIAL(B) hpmf_trial_kernel<B><<<grid, HT, 0, st>>>(v, d_coords, d_trial, d_mask, first_new, n_new, n_trials, \
            e1, m, c1, mc, h->d_carbon_ord, sa_env, w0, f0, v0, d_v, h->d_overflow)
        if (minb >= 8) LAUNCH_TRIAL(8); else if (minb >= 6) LAUNCH_TRIAL(6); else LAUNCH_TRIAL(4);
#undef LAUNCH_TRIAL
    }
    HCK(cudaGetLastError());
    int over = 0;
    HCK(cudaMemcpyAsync(&over, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    HCK(cudaStreamSynchronize(st));
    if (over) {
        HCK(cudaMemset(h->d_overflow, 0, sizeof(int)));
        h->state_level = 0;
        return mcsce_fail_shared("mcsce_hpmf_trials: more than 512 environment carbons change weight under one trial");
    }
    return 0;
}

extern "C" int mcsce_hpmf_carry(mcsce_hpmf h, int enable) {
    if (!h) return mcsce_fail_shared("mcsce_hpmf_carry: null handle");
    h->carry = enable != 0;
    h->state_level = 0;                     // the next mcsce_hpmf_trials call derives the environment from scratch
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
