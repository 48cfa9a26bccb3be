// mcsce_b200 - A7: hydrophobic potential of mean force on the Hasel-Hendrickson-Still surface, sm_100a.
//
//   sasa_kernel        SA_i = S_i * prod_{j != i, d_ij < R_i + R_j + 2 R_s} (1 - P_i P_ij b_ij / S_i)      (libsasa.py:89-108)
//   hpmf_pair_kernel   V = 1/2 sum_{i != j carbons, SA > A_c} tanh(SA_i) tanh(SA_j) sum_k H_k exp(-((r - C_k)/W_k)^2)
//                                                                                                          (libhpmf.py:92-109)
//   hpmf_finish_kernel fixed-order sum of the per-block partial sums
//
// One structure per blockIdx.y, 128 atoms (carbons) per block, partner atoms staged through shared memory in tiles
// of 128.  Numerics follow the reference: pair difference, squared norm and square root in float32 (np.linalg.norm
// of a float32 array: x*x, add.reduce over three elements, sqrt), everything after that in float64, the product
// over partners in ascending partner order.  Pairs that cannot pass the distance test are rejected on the float32
// squared distance; Gaussian terms beyond 18 A (< 2e-22 each) are dropped.

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

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

__global__ void __launch_bounds__(HT) sasa_kernel(HpmfDev s, const float *__restrict__ coords, double *__restrict__ sasa) {
    __shared__ float4 sp[HT];               // x, y, z, type (int bits): one 16-byte load per partner
    __shared__ double s_thr[MAX_SASA_TYPES * MAX_SASA_TYPES];
    __shared__ double s_R[MAX_SASA_TYPES], s_piRRs[MAX_SASA_TYPES];
    const int tid = threadIdx.x, nt = s.n_types;
    const long long base = (long long)blockIdx.y * s.n_atoms;
    for (int k = tid; k < nt * nt; k += HT) s_thr[k] = s.thr[k];
    if (tid < nt) { s_R[tid] = s.R[tid]; s_piRRs[tid] = s.piRRs[tid]; }
    const int i = blockIdx.x * HT + tid;
    const bool live = i < s.n_typed;
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
    for (int j0 = 0; j0 < s.n_typed; j0 += HT) {
        __syncthreads();
        const int jl = j0 + tid;
        if (jl < s.n_typed) {
            const float *p = coords + (base + s.typed_atom[jl]) * 3;
            sp[tid] = make_float4(p[0], p[1], p[2], __int_as_float(s.type_of[jl]));
        }
        __syncthreads();
        if (!live) continue;
        const int n = min(HT, s.n_typed - j0);
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
                    // dij + 1e-8 stays float32 in the reference (float32 array + Python scalar)
                    const double dd = (double)__fadd_rn(d, 1e-8f);
                    const double bij = __dmul_rn(__dmul_rn(s_piRRs[tj], __dsub_rn(thr, (double)d)),
                                                 __dadd_rn(1.0, __ddiv_rn(__dsub_rn(s_R[tj], Ri), dd)));
                    const double term = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(__dmul_rn(Pi, pij), bij), Si));
                    prod = __dmul_rn(prod, term);
                }
            }
        }
    }
    // an atom parked at 1e15 (not placed yet, the growth's convention) has no surface and no partners
    if (live) sasa[(long long)blockIdx.y * s.n_typed + i] = xi > ABSENT_X ? 0.0 : __dmul_rn(Si, prod);
}

__global__ void __launch_bounds__(HT) hpmf_pair_kernel(HpmfDev s, const float *__restrict__ coords,
                                                       const double *__restrict__ sasa, double *__restrict__ partial) {
    __shared__ float4 sp[HT];
    __shared__ double sw[HT];
    __shared__ double red[HT];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.y * s.n_atoms;
    const double *sa = sasa + (long long)blockIdx.y * s.n_typed;
    const int i = blockIdx.x * HT + tid;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    double wi = 0.0;
    if (i < s.n_carbon) {
        const int c = s.carbon[i];
        const float *p = coords + (base + s.typed_atom[c]) * 3;
        xi = p[0]; yi = p[1]; zi = p[2];
        const double a = sa[c];
        wi = (a > s.sa_threshold && xi <= ABSENT_X) ? tanh(a) : 0.0;
    }
    const double C0 = s.C[0], C1 = s.C[1], C2 = s.C[2], W0 = s.W[0], W1 = s.W[1], W2 = s.W[2];
    const double H0 = s.H[0], H1 = s.H[1], H2 = s.H[2];
    double acc = 0.0;
    for (int j0 = 0; j0 < s.n_carbon; j0 += HT) {
        __syncthreads();
        const int jl = j0 + tid;
        if (jl < s.n_carbon) {
            const int c = s.carbon[jl];
            const float *p = coords + (base + s.typed_atom[c]) * 3;
            const double a = sa[c];
            const double w = (a > s.sa_threshold && p[0] <= ABSENT_X) ? tanh(a) : 0.0;
            sw[tid] = w;
            sp[tid] = make_float4(p[0], p[1], p[2], w == 0.0 ? 0.f : 1.f);
        }
        __syncthreads();
        if (wi == 0.0) continue;
        const int n = min(HT, s.n_carbon - j0);
#pragma unroll 4
        for (int jj = 0; jj < n; ++jj) {
            const float4 q = sp[jj];
            if (q.w != 0.f && dist2_fast(xi, yi, zi, q.x, q.y, q.z) < HPMF_CUTOFF * HPMF_CUTOFF) {
                if (j0 + jj == i) continue;
                const double wj = sw[jj];
                const double r = (double)__fsqrt_rn(norm2_f32(xi, yi, zi, q.x, q.y, q.z));
                const double t0 = (r - C0) / W0, t1 = (r - C1) / W1, t2 = (r - C2) / W2;
                const double g = H0 * exp(-(t0 * t0)) + H1 * exp(-(t1 * t1)) + H2 * exp(-(t2 * t2));
                acc += wj * g;
            }
        }
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
// host side
// ------------------------------------------------------------------------------------------------
struct mcsce_hpmf_s {
    int device = 0;
    HpmfDev dev{};
    std::vector<void *> owned;
    double *d_partial = nullptr;
    size_t partial_cap = 0;
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
    for (int k = 0; k < 3; ++k) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    delete h;
    return 0;
}

static int launch_sasa(mcsce_hpmf h, int n_struct, const float *d_coords, double *d_sasa, cudaStream_t st) {
    const HpmfDev &v = h->dev;
    if (v.n_typed == 0) return 0;
    const int tiles = (v.n_typed + HT - 1) / HT;
    for (int c0 = 0; c0 < n_struct; c0 += 65535) {
        const int nc = n_struct - c0 < 65535 ? n_struct - c0 : 65535;
        sasa_kernel<<<dim3(tiles, nc), HT, 0, st>>>(v, d_coords + (size_t)c0 * v.n_atoms * 3, d_sasa + (size_t)c0 * v.n_typed);
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
                                                         d_sasa + (size_t)c0 * v.n_typed, h->d_partial + (size_t)c0 * tiles);
    }
    HCK(cudaGetLastError());
    HCK(cudaEventRecord(h->ev[2], st));
    hpmf_finish_kernel<<<(n_struct + 127) / 128, 128, 0, st>>>(n_struct, tiles, h->d_partial, d_v);
    HCK(cudaGetLastError());
    h->timed = true;
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
