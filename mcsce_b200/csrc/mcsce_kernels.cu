// mcsce_b200 — hand-written sm_100a kernels + C ABI for the MCSCE side-chain growth path.
//
// Kernels (one launch each per grown residue, all conformers at once):
//   place_trials_kernel      K1  chi angles -> placed trial side chains       (block = 32 trial rows, lane = one template atom)
//   score_special_kernel     K2s pairs within three bonds + intra-side-chain  (warp = 4 trials, exact float64)
//   score_generic_kernel<1>  K2  neighbour screen: clash test against the atoms near the residue; flags the trials that
//                                are +inf anyway                               (thread = R trial atoms, list tile in smem)
//   score_generic_kernel<0>  K2  trial atoms x environment atoms, generic pairs, over the trials the screen left
//                                                                               (thread = R trial atoms, env tiles in smem)
//   select_kernel            K3  clash mask + Boltzmann selection + append    (block = one conformer)
// plus init_env / input_energy / gather kernels.  Reference file:line citations are in
// include/mcsce_b200.h and DESIGN.md.
//
// Numerics contract: pair difference and d^2 in float32 with the reference's
// operation order and no FMA contraction (libcalc.py:64-67,73-76), so clash decisions are bit-exact;
// generic-pair energies in float32 with float64 accumulation every FLUSH pairs; special pairs and
// the selection in float64; chi rotations in float64 with float32 rounding where the reference rounds.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <functional>

#include "../../include/mcsce_b200.h"

// the hydrophobic PMF inside the growth (hpmf_kernels.cu)
int mcsce_hpmf_set_near_lists(mcsce_hpmf h, long long version, int n_res, const int *off, const int *atoms, const int *off2,
                              const int *atoms2);
double mcsce_hpmf_pair_range(mcsce_hpmf h);
long long mcsce_hpmf_near_version(mcsce_hpmf h);
int mcsce_hpmf_n_atoms(mcsce_hpmf h);
double mcsce_hpmf_sasa_range(mcsce_hpmf h);
int mcsce_hpmf_growth_begin(mcsce_hpmf h, int lane, const float *d_coords, int n_input, double *d_v_init, cudaStream_t st);
int mcsce_hpmf_growth_step(mcsce_hpmf h, int lane, int residue, int n_conf, const float *d_coords, int first_new, int n_new,
                           int n_trials, const float4 *d_trial4, long long trial4_stride, const uint8_t *d_mask, double *d_v,
                           int *d_overflow_conf, cudaStream_t st, const float *ca, float reach);

#define MAX_CLS MCSCE_MAX_CLS
#define MAX_SP_ENV MCSCE_MAX_SP_ENV
#define MAX_CHI MCSCE_MAX_CHI
#define MAX_ROT MCSCE_MAX_CHI_ROT
#define MAX_TMPL MCSCE_MAX_TMPL
#define MAX_SC MCSCE_MAX_SC

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return -1; }
// error text of the other translation units of the library (hpmf_kernels.cu)
int mcsce_fail_shared(const std::string &m) { return fail(m); }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// Coulomb prefactor: charges_ij *= 0.25; *= 1389.35  (libenergy.py:521-522)
#define COULOMB_PREF (0.25 * 1389.35)

// ------------------------------------------------------------------------------------------------
// device-side tables
// ------------------------------------------------------------------------------------------------
struct StepDev {
    int kind, type, n_sc, sc_start;
    int n_sp_env, sp_off, n_sp, pad0;
    int sp_env_slot[MAX_SP_ENV];
    // backbone-dependent
    int n_trials, trial_off, n_chi_lib, n_chi;
    float ca[4];
    double q1[4], q2[4];
    // side-chain atoms whose coordinates depend on the chi angles (var) and those that do not (fix: never in a
    // moving set, or the pivot of the first rotation - e.g. HA and CB): the latter are identical in every trial
    // of a conformer, so the growth scores them once per residue instead of once per trial
    int n_var, n_fix;
    unsigned char var_k[MAX_SC], fix_k[MAX_SC];
};

struct TmplDev {
    int n_atoms, n_chi;
    float xyz[MAX_TMPL][3];
    int sc_rank[MAX_TMPL];              // -1 for backbone atoms
    int axis_from[MAX_ROT], axis_to[MAX_ROT];
    unsigned move[MAX_ROT];
    double ori[MAX_ROT];
};

struct SysDev {
    int n_atoms, n_input, n_res, n_cls, n_types;
    const int *slot_of_atom, *cls_of_atom, *cls_start, *cls_active;
    const int *cls_prefix;              // [(n_res+1)*(n_cls+1)] exclusive prefix of cls_active per residue
    const double *charge;
    const float *cls_s2f, *cls_thr2f;
    const double *cls_e4, *cls_s2d, *cls_thrd;
    const double *cls_A, *cls_B;         // [n_cls*n_cls] e4*s2^6 and e4*s2^3: A/d^12 - B/d^6
    const StepDev *steps;
    const TmplDev *tmpl;
    const int *sp_k, *sp_partner;
    const double *sp_coef;
    const double *sp_A, *sp_B, *sp_Q;   // the same coefficients, one array each (coalesced loads in the special-pair kernel)
    const float *sp_thr2f;
    const int *resnum;
    int n_bb_pairs;
    const int *bb_i, *bb_j;
    const double *bb_coef;
    // backbone
    const double *chi_mean, *chi_sigma, *prob;
    const float *input_xyz, *rigid_xyz;
    int n_trials_total;
    // neighbour screen (per backbone): for every grown residue the active environment slots that can come near its
    // side chain, class-major; scr_prefix[step*(n_cls+1)+c] = start of class c inside the residue's list
    const int *scr_off, *scr_slot, *scr_prefix;
};

// Scalars that change from call to call live in device memory so that a captured CUDA graph of a growth can be
// replayed with a new seed / conformer offset / temperature without re-capturing.
struct CallParams {
    unsigned long long seed;
    long long conf_id0;
    double beta, noise_deg;
};

__global__ void set_call_params_kernel(CallParams *dst, CallParams v) { *dst = v; }

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dist2_ref(float ax, float ay, float az, float bx, float by, float bz) {
    // x = b - a ; x*x + y*y + z*z, float32, unfused (libcalc.py:64-67)
    float x = __fsub_rn(bx, ax), y = __fsub_rn(by, ay), z = __fsub_rn(bz, az);
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__device__ __forceinline__ void quat_rotate(const double b1, const double b2, const double b3, const double b4,
                                            double x, double y, double z, double &ox, double &oy, double &oz) {
    // two Hamilton products evaluated term by term, left to right, unfused (libcalc.py:359-366, 404-412).  The point is a pure
    // quaternion (0, x, y, z): the four products with its zero scalar part are +-0 and adding them changes nothing but the
    // sign of an exact zero, so they are left out (8 of 49 float64 operations; K1 is float64-issue bound)
    double c1 = __dsub_rn(__dsub_rn(-__dmul_rn(b2, x), __dmul_rn(b3, y)), __dmul_rn(b4, z));
    double c2 = __dsub_rn(__dadd_rn(__dmul_rn(b1, x), __dmul_rn(b3, z)), __dmul_rn(b4, y));
    double c3 = __dadd_rn(__dsub_rn(__dmul_rn(b1, y), __dmul_rn(b2, z)), __dmul_rn(b4, x));
    double c4 = __dsub_rn(__dadd_rn(__dmul_rn(b1, z), __dmul_rn(b2, y)), __dmul_rn(b3, x));
    double n2 = -b2, n3 = -b3, n4 = -b4;
    ox = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(c1, n2), __dmul_rn(c2, b1)), __dmul_rn(c3, n4)), __dmul_rn(c4, n3));
    oy = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(c1, n3), __dmul_rn(c2, n4)), __dmul_rn(c3, b1)), __dmul_rn(c4, n2));
    oz = __dadd_rn(__dsub_rn(__dadd_rn(__dmul_rn(c1, n4), __dmul_rn(c2, n3)), __dmul_rn(c3, n2)), __dmul_rn(c4, b1));
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: one call per (conformer, residue, trial, chi)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ double u01_open(unsigned a, unsigned b) {   // (0,1]
    unsigned long long v = ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11);
    return ((double)(v & ((1ull << 53) - 1)) + 1.0) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u01_halfopen(unsigned a, unsigned b) {   // [0,1)
    unsigned long long v = ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11);
    return (double)(v & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
}

// half-angle sine/cosine of chi rotation k: angle = wrap(template dihedral - target)   (libscbuild.py:44-54)
__device__ __forceinline__ void chi_half_sincos(double ori, double chi_deg, double &sn, double &cs) {
    double rot = ori - chi_deg * CUDART_PI / 180.0;                   // measured - target (libscbuild.py:47)
    if (rot > CUDART_PI) rot -= 2.0 * CUDART_PI;
    else if (rot < -CUDART_PI) rot += 2.0 * CUDART_PI;
    sincos(rot * 0.5, &sn, &cs);
}

// chi rotations of one trial held one template atom per lane; float32 storage after every rotation
// (libscbuild.py:108).  Rotation k turns the atoms of the moving set about MINUS the unit axis through the
// pivot atom; sn[k], cs[k] = sin, cos of half its angle (warp-uniform).
__device__ __forceinline__ void apply_chi_rotations(const TmplDev &tm, int n_rot, const double *sn_k, const double *cs_k,
                                                    int lane, bool has, float &x, float &y, float &z) {
    for (int k = 0; k < n_rot; ++k) {
        const int af = tm.axis_from[k], at = tm.axis_to[k];
        const float fx = __shfl_sync(0xffffffffu, x, af), fy = __shfl_sync(0xffffffffu, y, af), fz = __shfl_sync(0xffffffffu, z, af);
        const float px = __shfl_sync(0xffffffffu, x, at), py = __shfl_sync(0xffffffffu, y, at), pz = __shfl_sync(0xffffffffu, z, at);
        // unit axis in float32: (to - from)/|to - from|   (libscbuild.py:105,115,169,185,199)
        const float ux = __fsub_rn(px, fx), uy = __fsub_rn(py, fy), uz = __fsub_rn(pz, fz);
        // np.linalg.norm of a float32 3-vector = sqrt(x.dot(x)) through OpenBLAS sdot's scalar tail: float32 products summed
        // in a double, the sum rounded to float32, float32 sqrt (checked against NumPy: identical on 20000 random vectors)
        const double acc = __dadd_rn(__dadd_rn((double)__fmul_rn(ux, ux), (double)__fmul_rn(uy, uy)), (double)__fmul_rn(uz, uz));
        const float nrm = __fsqrt_rn(__double2float_rn(acc));
        const float ex = __fdiv_rn(ux, nrm), ey = __fdiv_rn(uy, nrm), ez = __fdiv_rn(uz, nrm);
        const double sn = sn_k[k], cs = cs_k[k];
        const double b2 = sn * (double)(-ex), b3 = sn * (double)(-ey), b4 = sn * (double)(-ez);
        if (has && ((tm.move[k] >> lane) & 1u)) {
            const float dx = __fsub_rn(x, px), dy = __fsub_rn(y, py), dz = __fsub_rn(z, pz);   // float32 (libscbuild.py:44)
            double ox, oy, oz;
            quat_rotate(cs, b2, b3, b4, (double)dx, (double)dy, (double)dz, ox, oy, oz);
            x = __double2float_rn(__dadd_rn(ox, (double)px));
            y = __double2float_rn(__dadd_rn(oy, (double)py));
            z = __double2float_rn(__dadd_rn(oz, (double)pz));
        }
    }
}

// stand-alone template rotation (C-ABI mcsce_rotate_template): all template atoms, template frame, no placement
__global__ void __launch_bounds__(128) rotate_template_kernel(TmplDev tm, int n_rot, int n_rows, const double *chis, float *out) {
    __shared__ double s_sn[4][MAX_ROT], s_cs[4][MAX_ROT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * 4 + warp;
    if (t >= n_rows) return;
    const bool has = lane < tm.n_atoms;
    float x = 0.f, y = 0.f, z = 0.f;
    if (has) { x = tm.xyz[lane][0]; y = tm.xyz[lane][1]; z = tm.xyz[lane][2]; }
    if (lane < n_rot) chi_half_sincos(tm.ori[lane], chis[(long long)t * MAX_CHI + lane], s_sn[warp][lane], s_cs[warp][lane]);
    __syncwarp();
    apply_chi_rotations(tm, n_rot, s_sn[warp], s_cs[warp], lane, has, x, y, z);
    if (has) { float *o = out + ((long long)t * tm.n_atoms + lane) * 3; o[0] = x; o[1] = y; o[2] = z; }
}

// stand-alone rigid placement (C-ABI mcsce_place_template): float32 in, float64 out (libcalc.py:465,480-482)
__global__ void place_template_kernel(int n, const float *xyz, double q10, double q11, double q12, double q13, double q20,
                                      double q21, double q22, double q23, float cax, float cay, float caz, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ox, oy, oz;
    quat_rotate(q10, q11, q12, q13, (double)xyz[i * 3], (double)xyz[i * 3 + 1], (double)xyz[i * 3 + 2], ox, oy, oz);
    const float rx = __double2float_rn(ox), ry = __double2float_rn(oy), rz = __double2float_rn(oz);
    quat_rotate(q20, q21, q22, q23, (double)rx, (double)ry, (double)rz, ox, oy, oz);
    out[i * 3] = __dadd_rn(ox, (double)cax); out[i * 3 + 1] = __dadd_rn(oy, (double)cay); out[i * 3 + 2] = __dadd_rn(oz, (double)caz);
}

// ------------------------------------------------------------------------------------------------
// K1: chi -> xyz.   grid (ceil(n_trials/K1_TRIALS), n_sets), block 128: K1_TRIALS trial rows per block
// ------------------------------------------------------------------------------------------------
struct PlaceArgs {
    SysDev s;
    int step;
    int n_trials;             // rows per set
    const double *chis_in;    // explicit rows [set][row_stride] or null
    long long chis_set_stride;  // doubles between sets in chis_in / chis_out
    long long chis_row_off;     // offset (in rows) of this residue's first row inside a set
    double *chis_out;         // optional dump, same layout
    const double *ori_in;     // optional replay input [set][row][MAX_ROT]: the dihedral the reference measured before each rotation (rad);
                              // NaN or null = the template constant
    float4 *trial;            // [set][n_trials*n_sc]
    long long trial_set_stride;
    const unsigned char *alive;
    const CallParams *cp;     // seed, conformer offset, perturbation sigma (device memory)
};

#define K1_TRIALS 32          // trial rows per block: the scalar work per (row, chi) - Philox, Box-Muller, the half-angle
                              // sincos - is done one thread per (row, chi) first, then each warp rotates/places 8 rows
__global__ void __launch_bounds__(128) place_trials_kernel(PlaceArgs a) {
    __shared__ double s_sn[K1_TRIALS][MAX_ROT], s_cs[K1_TRIALS][MAX_ROT];
    const int set = blockIdx.y;
    if (a.alive && !a.alive[set]) return;
    const int t0 = blockIdx.x * K1_TRIALS;
    const StepDev &st = a.s.steps[a.step];
    const TmplDev &tm = a.s.tmpl[st.type];

    // --- phase 1: trial chi rows (degrees) and the rotations' half-angle sin/cos, one thread per (row, chi column) ---
    for (int idx = threadIdx.x; idx < K1_TRIALS * 8; idx += 128) {
        const int tl = idx >> 3, c = idx & 7, t = t0 + tl;
        if (t >= a.n_trials || c >= MAX_CHI) continue;
        const long long row = a.chis_row_off + t;
        const long long at = (long long)set * a.chis_set_stride + row * MAX_CHI + c;
        double v = 0.0;
        if (a.chis_in) {
            v = a.chis_in[at];
            if (a.chis_out) a.chis_out[at] = v;
        } else {
            // wrap(mean + sigma*z1) + noise*z2   (rotamer_library.py:155,165 ; side_chain_builder.py:173)
            if (c < st.n_chi_lib) {
                const long long trow = (long long)st.trial_off + t;
                const long long cid = a.cp->conf_id0 + set;
                const unsigned long long seed = a.cp->seed;
                uint4 r = philox4x32(make_uint4((unsigned)cid, (unsigned)(cid >> 32),
                                                (unsigned)a.step, ((unsigned)t << 3) | (unsigned)c),
                                     make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
                double u1 = u01_open(r.x, r.y), u2 = u01_halfopen(r.z, r.w);
                double rad = sqrt(-2.0 * log(u1)), sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
                v = a.s.chi_mean[trow * MAX_CHI + c];
                const double sg = a.s.chi_sigma[trow * MAX_CHI + c];
                if (sg != 0.0) {
                    v = v + sg * (rad * cs);
                    if (v > 180.0) v -= 360.0;
                    if (v < -180.0) v += 360.0;
                }
                v = v + a.cp->noise_deg * (rad * sn);
            }
            if (a.chis_out) a.chis_out[at] = v;
        }
        if (c < st.n_chi) {
            // The reference re-measures the dihedral in float32 before every rotation (libscbuild.py:113,167,183,197); that
            // value is the template constant plus float32 noise of the host's NumPy (DESIGN "K1 fidelity").  Parity runs feed
            // the recorded values back in; otherwise the constant is used.
            double ori = tm.ori[c];
            if (a.ori_in) {
                const double m = a.ori_in[((long long)set * a.chis_set_stride / MAX_CHI + row) * MAX_ROT + c];
                if (m == m) ori = m;
            }
            chi_half_sincos(ori, v, s_sn[tl][c], s_cs[tl][c]);
        }
    }
    __syncthreads();

    // --- phase 2: one template atom per lane; each warp takes K1_TRIALS/4 rows ---
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool has = lane < tm.n_atoms;
    float x0 = 0.f, y0 = 0.f, z0 = 0.f;
    if (has) { x0 = tm.xyz[lane][0]; y0 = tm.xyz[lane][1]; z0 = tm.xyz[lane][2]; }
    const int rank = has ? tm.sc_rank[lane] : -1;
    const int n_rot = st.n_chi;
    float4 *out = a.trial + (long long)set * a.trial_set_stride;
    for (int j = 0; j < K1_TRIALS / 4; ++j) {
        const int tl = warp * (K1_TRIALS / 4) + j, t = t0 + tl;
        if (t >= a.n_trials) break;
        float x = x0, y = y0, z = z0;
        apply_chi_rotations(tm, n_rot, s_sn[tl], s_cs[tl], lane, has, x, y, z);
        // rigid placement: two rotations (first stored to float32) + CA (libcalc.py:465,480-482)
        if (rank >= 0) {
            double ox, oy, oz;
            quat_rotate(st.q1[0], st.q1[1], st.q1[2], st.q1[3], (double)x, (double)y, (double)z, ox, oy, oz);
            const float rx = __double2float_rn(ox), ry = __double2float_rn(oy), rz = __double2float_rn(oz);
            quat_rotate(st.q2[0], st.q2[1], st.q2[2], st.q2[3], (double)rx, (double)ry, (double)rz, ox, oy, oz);
            float4 o;
            o.x = __double2float_rn(__dadd_rn(ox, (double)st.ca[0]));
            o.y = __double2float_rn(__dadd_rn(oy, (double)st.ca[1]));
            o.z = __double2float_rn(__dadd_rn(oz, (double)st.ca[2]));
            o.w = 0.f;
            out[(long long)t * st.n_sc + rank] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: generic pairs.  grid (trial tiles, n_sets, env_split), block SCORE_THREADS.
// ------------------------------------------------------------------------------------------------
#ifndef SCORE_THREADS
#define SCORE_THREADS 128               // 128 x 2 measured best on B200 (tools/time_variants.py): finer CTA granularity per launch
#endif
#ifndef SCORE_R
#define SCORE_R 2                       // trial atoms per thread at most: a tile holds up to SCORE_R rows of 128 atoms (see tile_layout;
                                        // 3 compiles to 127 registers without spills and was measured: no gain)
#endif
#ifndef SCORE_MINB
#define SCORE_MINB 4                    // min resident blocks per SM asked of the compiler (keeps the kernel at <= 128 registers: 4 CTAs/SM)
#endif
#define TILE_A (SCORE_THREADS * SCORE_R)
#define SCREEN_TILE_A (SCORE_THREADS * 2)   // the clash-only screen pass keeps 256-atom tiles (it is overhead bound, not pair bound)
#ifndef TILE_E
#define TILE_E 1024                     // environment atoms staged per smem tile
#endif
#ifndef NO_RECIP_CORRECTION
#define NO_RECIP_CORRECTION 1           // 1: 1/d2 = y*y of the raw rsqrt; 0: one Newton step on the reciprocal (2 more FMA-pipe ops per
                                        // pair, -12 % speed).  Measured K2 error vs the reference: worst trial 0.21x (1) / 0.15x (0) of the tolerance
#endif
#ifndef PAIR_UNROLL
#define PAIR_UNROLL 4                   // environment-atom pairs per unrolled iteration of the inner loop
#endif
constexpr int kPairUnroll = PAIR_UNROLL;
#ifndef RCP_MIX
#define RCP_MIX 0                       // 1: alternate Newton-corrected y*y with MUFU.RCP for 1/d2
#endif
#ifndef FLUSH
#define FLUSH 32                        // pairs accumulated in float32 before a float64 flush
#endif

// Tile layout of the scoring pass for one conformer: nk trials of n_per atoms each (+ n_fix chi-independent atoms riding in the
// last tile) are cut into n_tiles tiles of tpt whole trials; a tile runs on one CTA with up to max_rows atoms per thread (rows of
// SCORE_THREADS atoms).  Two rules: packed - fill every tile to the cap, remainder last (the shipped one, 2 rows); balanced -
// the number of tiles that minimises  tiles x (rows + overhead).  Measured on config 3 x 2048 (round 2, tools/strong_curve.py):
// packed/2 rows 543 ms, packed/3 rows 550, balanced/3 rows 570, balanced/2 rows 578, packed/1 row 602 - a half-empty last tile
// costs far less than its idle lanes suggest, because the other CTAs resident on the SM take the issue slots it leaves free.
struct TileRule { int max_rows, packed; float overhead; };
__host__ __device__ __forceinline__ void tile_layout(TileRule rule, int nk, int n_per, int n_fix, int &n_tiles, int &tpt) {
    n_tiles = 1; tpt = nk > 0 ? nk : 1;
    if (nk <= 0) return;
    const int cap = SCORE_THREADS * rule.max_rows;
    const int atoms = nk * n_per + n_fix;
    if (rule.packed) {       // fill tiles to the cap, the remainder (and the chi-independent atoms) in the last one
        tpt = cap / n_per; if (tpt < 1) tpt = 1;
        n_tiles = (nk + tpt - 1) / tpt;
        if ((nk - (n_tiles - 1) * tpt) * n_per + n_fix > cap) ++n_tiles;    // (an extra tile for the chi-independent atoms)
        return;
    }
    const int n_min = (atoms + cap - 1) / cap;
    float best = 1.0e30f;
    for (int n = (n_min > 0 ? n_min : 1); n <= n_min + 5; ++n) {
        const int t = (nk + n - 1) / n;
        const int tile_atoms = t * n_per + n_fix;                 // the last tile carries the chi-independent atoms: bound every tile by it
        if (tile_atoms > cap) { if (t <= 1) break; continue; }
        const int used = (nk + t - 1) / t;                        // tiles that actually hold trials
        const float cost = (float)used * ((float)((tile_atoms + SCORE_THREADS - 1) / SCORE_THREADS) + rule.overhead);
        if (cost < best) { best = cost; n_tiles = used; tpt = t; }
        if (t <= 1) break;
    }
}

struct ScoreArgs {
    SysDev s;
    int step;
    int n_trials;                      // per set
    int trials_per_tile;
    int n_split;
    const float4 *env;                 // [set][n_atoms]  slot order
    const float4 *trial;               // [set][n_trials*n_sc]
    long long trial_set_stride;
    double2 *partial;                  // [set][n_trials + dedup][n_split] (energy, clash); slot n_trials = chi-independent atoms
    const unsigned char *alive;
    int dedup;                         // 1: score chi-independent atoms once per residue (trial 0's copy); they ride in the last tile
    TileRule rule;                     // how the scoring pass cuts a conformer's trials into tiles
    unsigned char *screen;             // [set][n_trials] neighbour screen: 1 = the trial clashes with a nearby atom.  Written by
                                       // the screen pass; the main pass (null = no screen) scores only the trials left
    const double2 *sp_partial;         // screen pass: the special-pair kernel's results [set][n_trials]; its clash flags are merged in
    int *n_keep;                       // [set] trials the screen left (added up by the screen pass, zeroed again by the selection)
};

// Inner loop over one class segment of the staged tile for RR trial atoms of this thread, two
// environment atoms per iteration with Blackwell's packed-pair FP32 instructions (FADD2/FMUL2/FFMA2,
// PTX add/mul/fma.f32x2, sm_100+): the FMA pipe does the same number of lane-operations but takes half
// the issue slots, which is what bounds this kernel.  Per environment-atom pair and trial atom:
//   d   = e - a                                   3 FADD2
//   d2  = fma(dz,dz,fma(dy,dy,dx*dx))             FMUL2 + 2 FFMA2   (fused: <= 4 ulp from the reference's unfused d2)
//   dmin= min(dmin, d2.x, d2.y)                   clash decision deferred to the end of the segment
//   y   = rsqrt.approx(d2)                        2 MUFU (max rel. error 2^-22.9): used as is for q/d
//   w   = y*y                                     FMUL2: 1/d2 to 3.2e-7, i.e. d^-12 to 2.2e-6 worst case, 0.22x the 1e-5 tolerance
//                                                 (NO_RECIP_CORRECTION=0 adds the Newton step w += w*(1 - d2*w): 2 FFMA2)
//   i3  = w*w*w ; a6 += i3*i3 ; a3 += i3          2 FMUL2 + FFMA2 + FADD2: sum d^-12, sum d^-6 (class constants at the flush)
//   cq += q*y                                     FFMA2
// The staged tile is stored pair-interleaved (x0 x1 y0 y1 | z0 z1 q0 q1) and every class segment is padded
// to an even length with a parked dummy atom, so the loop has no edge handling.
// The clash test must reproduce  sqrt(float64(d2_ref)) < thr  exactly (d2_ref = the reference's unfused
// float32 sum, libcalc.py:64-67).  thr2 is the float32 value with  d2_ref < thr2  <=>  that predicate, so
// the fused minimum decides unless it lands within 1e-6 relative of thr2; then the segment is re-scanned
// with the reference's operation order (practically never).
template <int RR>
__device__ __forceinline__ void score_segment(const float4 *__restrict__ s_env, int p0, int p1,
                                              const float2 (&nx)[SCORE_R], const float2 (&ny)[SCORE_R], const float2 (&nz)[SCORE_R],
                                              const float (&thr2)[SCORE_R], double (&E6)[SCORE_R], double (&E3)[SCORE_R],
                                              double (&EC)[SCORE_R], bool (&clash)[SCORE_R]) {
    float dmin[RR];
#pragma unroll
    for (int r = 0; r < RR; ++r) dmin[r] = CUDART_INF_F;
    const float2 one2 = make_float2(1.0f, 1.0f);
    for (int pb = p0; pb < p1; pb += FLUSH / 2) {
        const int pe = min(pb + FLUSH / 2, p1);
        float2 a6[RR], a3[RR], cq[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) { a6[r] = make_float2(0.f, 0.f); a3[r] = a6[r]; cq[r] = a6[r]; }
#pragma unroll kPairUnroll
        for (int p = pb; p < pe; ++p) {
            const float4 exy = s_env[2 * p], ezq = s_env[2 * p + 1];
            const float2 ex = make_float2(exy.x, exy.y), ey = make_float2(exy.z, exy.w);
            const float2 ez = make_float2(ezq.x, ezq.y), eq = make_float2(ezq.z, ezq.w);
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const float2 dx = __fadd2_rn(ex, nx[r]), dy = __fadd2_rn(ey, ny[r]), dz = __fadd2_rn(ez, nz[r]);
                const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                dmin[r] = fminf(dmin[r], fminf(d2.x, d2.y));
                float2 y;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(d2.x));
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(d2.y));
                float2 w;
#if RCP_MIX
                if (p & 1) {      // every other pair: 1/d2 straight from the MUFU (<= 1 ulp), balancing the FMA and XU pipes
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w.x) : "f"(d2.x));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w.y) : "f"(d2.y));
                } else
#endif
                {
                    w = __fmul2_rn(y, y);
#if !NO_RECIP_CORRECTION
                    const float2 nd2 = make_float2(-d2.x, -d2.y);
                    w = __ffma2_rn(w, __ffma2_rn(nd2, w, one2), w);
#endif
                }
                const float2 i3 = __fmul2_rn(__fmul2_rn(w, w), w);
                a6[r] = __ffma2_rn(i3, i3, a6[r]);
                a3[r] = __fadd2_rn(a3[r], i3);
                cq[r] = __ffma2_rn(eq, y, cq[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            E6[r] += (double)(a6[r].x + a6[r].y); E3[r] += (double)(a3[r].x + a3[r].y); EC[r] += (double)(cq[r].x + cq[r].y);
        }
    }
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        if (dmin[r] < thr2[r] * (1.0f - 1.0e-6f)) clash[r] = true;
        else if (dmin[r] < thr2[r] * (1.0f + 1.0e-6f)) {
            for (int p = p0; p < p1; ++p) {
                const float4 exy = s_env[2 * p], ezq = s_env[2 * p + 1];
                clash[r] |= dist2_ref(-nx[r].x, -ny[r].x, -nz[r].x, exy.x, exy.z, ezq.x) < thr2[r];
                clash[r] |= dist2_ref(-nx[r].x, -ny[r].x, -nz[r].x, exy.y, exy.w, ezq.y) < thr2[r];
            }
        }
    }
}

template <int RR>
__device__ __forceinline__ void score_tile(const ScoreArgs &a, const float4 *__restrict__ s_env, const int *s_seg, int c_first,
                                           int c_last, const float2 (&nx)[SCORE_R], const float2 (&ny)[SCORE_R],
                                           const float2 (&nz)[SCORE_R], const int (&ci)[SCORE_R], double (&e_lj)[SCORE_R],
                                           double (&e_c)[SCORE_R], bool (&clash)[SCORE_R]) {
    const int n_cls = a.s.n_cls;
    for (int c = c_first; c <= c_last; ++c) {
        const int p0 = s_seg[c] >> 1, p1 = s_seg[c + 1] >> 1;          // padded, pair units
        if (p0 >= p1) continue;
        float thr2[SCORE_R];
        double A[SCORE_R], B[SCORE_R];
        double E6[SCORE_R], E3[SCORE_R], EC[SCORE_R];
        // class-pair constants, loaded before the pair loop so that their latency hides under it:
        // A/d^12 - B/d^6 with A = e4*s2^6, B = e4*s2^3 (libenergy.py:478-481), float64, products formed on the host
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int idx = ci[r] * n_cls + c;
            thr2[r] = __ldg(&a.s.cls_thr2f[idx]); A[r] = __ldg(&a.s.cls_A[idx]); B[r] = __ldg(&a.s.cls_B[idx]);
            E6[r] = 0.0; E3[r] = 0.0; EC[r] = 0.0;
        }
        score_segment<RR>(s_env, p0, p1, nx, ny, nz, thr2, E6, E3, EC, clash);
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            e_lj[r] += A[r] * E6[r] - B[r] * E3[r];
            e_c[r] += EC[r];
        }
    }
}

// Clash-only variants for the neighbour screen: distances and the running minimum, same decision rule.
template <int RR>
__device__ __forceinline__ void clash_segment(const float4 *__restrict__ s_env, int p0, int p1,
                                              const float2 (&nx)[SCORE_R], const float2 (&ny)[SCORE_R], const float2 (&nz)[SCORE_R],
                                              const float (&thr2)[SCORE_R], bool (&clash)[SCORE_R]) {
    float dmin[RR];
#pragma unroll
    for (int r = 0; r < RR; ++r) dmin[r] = CUDART_INF_F;
#pragma unroll kPairUnroll
    for (int p = p0; p < p1; ++p) {
        const float4 exy = s_env[2 * p], ezq = s_env[2 * p + 1];
        const float2 ex = make_float2(exy.x, exy.y), ey = make_float2(exy.z, exy.w), ez = make_float2(ezq.x, ezq.y);
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const float2 dx = __fadd2_rn(ex, nx[r]), dy = __fadd2_rn(ey, ny[r]), dz = __fadd2_rn(ez, nz[r]);
            const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            dmin[r] = fminf(dmin[r], fminf(d2.x, d2.y));
        }
    }
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        if (dmin[r] < thr2[r] * (1.0f - 1.0e-6f)) clash[r] = true;
        else if (dmin[r] < thr2[r] * (1.0f + 1.0e-6f)) {
            for (int p = p0; p < p1; ++p) {
                const float4 exy = s_env[2 * p], ezq = s_env[2 * p + 1];
                clash[r] |= dist2_ref(-nx[r].x, -ny[r].x, -nz[r].x, exy.x, exy.z, ezq.x) < thr2[r];
                clash[r] |= dist2_ref(-nx[r].x, -ny[r].x, -nz[r].x, exy.y, exy.w, ezq.y) < thr2[r];
            }
        }
    }
}

template <int RR>
__device__ __forceinline__ void clash_tile(const ScoreArgs &a, const float4 *__restrict__ s_env, const int *s_seg, int c_first,
                                           int c_last, const float2 (&nx)[SCORE_R], const float2 (&ny)[SCORE_R],
                                           const float2 (&nz)[SCORE_R], const int (&ci)[SCORE_R], bool (&clash)[SCORE_R]) {
    const int n_cls = a.s.n_cls;
    for (int c = c_first; c <= c_last; ++c) {
        const int p0 = s_seg[c] >> 1, p1 = s_seg[c + 1] >> 1;
        if (p0 >= p1) continue;
        float thr2[SCORE_R];
#pragma unroll
        for (int r = 0; r < RR; ++r) thr2[r] = __ldg(&a.s.cls_thr2f[ci[r] * n_cls + c]);
        clash_segment<RR>(s_env, p0, p1, nx, ny, nz, thr2, clash);
    }
}

#define TILE_SLOTS (TILE_E + 2 * MAX_CLS)      // staged atoms + one pad per class, even
#define SCREEN_WORDS 64                         // trials per residue the screen can index (32 per word)

// SCREEN = false: the scoring pass (energies + clash against every active environment atom), over all trials or - with a
//                 screen - over the trials the screen left, addressed through a compacted index.
// SCREEN = true : the neighbour screen, clash test only against the residue's short list of nearby environment atoms
//                 (grid (trial tiles, n_sets, 1)); a trial flagged here is +inf whatever else it touches, so the
//                 scoring pass skips it.  A clash the list misses is still found by the scoring pass: the list only
//                 has to be likely, not complete.
template <bool SCREEN>
__global__ void __launch_bounds__(SCORE_THREADS, SCORE_MINB) score_generic_kernel(ScoreArgs a) {
    __shared__ float4 s_env2[2][TILE_SLOTS / 2 * 2];  // two tile buffers; pair p: [2p] = x0 x1 y0 y1, [2p+1] = z0 z1 q0 q1
    __shared__ int s_prefix[MAX_CLS + 1];
    __shared__ int s_seg3[3][MAX_CLS + 1];            // padded start of every class segment inside the tile: tiles t, t+1, t+2
    __shared__ int s_chunk3[3][MAX_CLS + 1];          // 32-atom staging chunks before every class segment
    __shared__ int s_sp[MAX_SP_ENV], s_sp_c[MAX_SP_ENV];
    __shared__ double s_e[SCREEN ? 1 : TILE_A];
    __shared__ unsigned char s_clash[TILE_A];
    __shared__ unsigned s_mask[SCREEN ? 1 : SCREEN_WORDS];     // surviving trials, one bit each
    __shared__ int s_cnt[SCREEN ? 2 : SCREEN_WORDS + 1];       // exclusive prefix of their counts per word

    const int set = blockIdx.y;
    if (a.alive && !a.alive[set]) return;
    const bool compact = !SCREEN && a.screen != nullptr;
    const int tid = threadIdx.x;
    const StepDev &st = a.s.steps[a.step];
    const int n_cls = a.s.n_cls;
    const int n_sc = st.n_sc;
    // Tiles of the scoring pass: chosen per conformer from the number of trials the screen left (tile_layout); the
    // chi-independent atoms ride in the last tile, tiles beyond have nothing to do.
    const int n_per = a.dedup ? st.n_var : n_sc;                 // atoms per trial in a tile
    int lay_tiles = gridDim.x, lay_tpt = a.trials_per_tile;
    if (!SCREEN) {
        tile_layout(a.rule, compact ? a.n_keep[set] : a.n_trials, n_per, a.dedup ? st.n_fix : 0, lay_tiles, lay_tpt);
        if ((int)blockIdx.x >= lay_tiles) return;
    }

    const int warp = tid >> 5, lane = tid & 31;
    {   // start of every class in the order the pass walks: the residue's screen list, or the active environment
        const int *src = (SCREEN ? a.s.scr_prefix : a.s.cls_prefix) + a.step * (n_cls + 1);
        for (int c = tid; c <= n_cls; c += SCORE_THREADS) s_prefix[c] = src[c];
    }
    __syncthreads();
    if (!SCREEN && tid < MAX_SP_ENV) {      // special environment atoms: where they sit in that order (-1: not active)
        int ae = -1, cc = 0;
        if (tid < st.n_sp_env) {
            const int slot = st.sp_env_slot[tid];
            int c_lo = 0, c_hi = n_cls;       // largest c with cls_start[c] <= slot
            while (c_hi - c_lo > 1) { int m = (c_lo + c_hi) >> 1; if (a.s.cls_start[m] <= slot) c_lo = m; else c_hi = m; }
            const int r = slot - a.s.cls_start[c_lo];
            if (r < s_prefix[c_lo + 1] - s_prefix[c_lo]) { ae = s_prefix[c_lo] + r; cc = c_lo; }
        }
        s_sp[tid] = ae; s_sp_c[tid] = cc;
    }
    int n_keep = a.n_trials;
    if (!SCREEN) {
        if (compact) {
            const unsigned char *flag = a.screen + (long long)set * a.n_trials;
            const int n_words = (a.n_trials + 31) >> 5;
            for (int wb = 0; wb < n_words; wb += SCORE_THREADS / 32) {
                const int w = wb + (tid >> 5), t = w * 32 + (tid & 31);
                const unsigned m = __ballot_sync(0xffffffffu, t < a.n_trials && flag[t] == 0);
                if ((tid & 31) == 0 && w < n_words) s_mask[w] = m;
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int w = 0; w < n_words; ++w) { s_cnt[w] = acc; acc += __popc(s_mask[w]); }
                s_cnt[n_words] = acc;
            }
            __syncthreads();
            n_keep = s_cnt[n_words];
        }
    }
    __syncthreads();
    const int n_active = s_prefix[n_cls];
    // j-th trial this pass scores -> its index among the residue's trials
    auto trial_of = [&](int j) -> int {
        if (SCREEN || !compact) return j;
        int w = 0;
        while (s_cnt[w + 1] <= j) ++w;
        return w * 32 + (int)__fns(s_mask[w], 0, j - s_cnt[w] + 1);
    };

    // this block's trial atoms: atom la = r*SCORE_THREADS + tid of the tile.  With dedup the tiles hold only the
    // chi-dependent atoms of whole trials; the chi-independent atoms (trial 0's copy) ride in the free lanes of the last tile.
    const bool last = !SCREEN && (int)blockIdx.x == lay_tiles - 1;
    const int t0 = blockIdx.x * lay_tpt;
    const int nt = max(0, min(lay_tpt, n_keep - t0));
    const int n_varatoms = nt * n_per;
    const int n_fixatoms = (!SCREEN && a.dedup && last) ? st.n_fix : 0;
    const int n_local = n_varatoms + n_fixatoms;
    if (n_local == 0) return;                                    // every trial of this tile was screened out
    const float4 *trial0 = a.trial + (long long)set * a.trial_set_stride;

    // Atom slots are row-major, atom la = r*SCORE_THREADS + tid, so a short tile keeps every warp at the same (smaller)
    // number of rows.  (Handing the idle warps of a short tile shares of the environment was measured twice - per class
    // segment and per quarter of the staged tile - and gained 0-1.5 %: not kept.)
    float2 nx[SCORE_R], ny[SCORE_R], nz[SCORE_R];     // minus the trial atom, duplicated into both halves
    int ci[SCORE_R], katom[SCORE_R];
    double e_lj[SCORE_R], e_c[SCORE_R];
    bool clash[SCORE_R];
#pragma unroll
    for (int r = 0; r < SCORE_R; ++r) {
        const int la = r * SCORE_THREADS + tid;
        int k = 0; float4 p = trial0[0];
        if (la < n_varatoms) {
            const int tl = la / n_per, kk = la % n_per;
            k = a.dedup ? st.var_k[kk] : kk;
            p = trial0[(long long)trial_of(t0 + tl) * n_sc + k];
        } else if (la < n_local) {
            k = st.fix_k[la - n_varatoms];
            p = trial0[k];
        }
        katom[r] = k;
        nx[r] = make_float2(-p.x, -p.x); ny[r] = make_float2(-p.y, -p.y); nz[r] = make_float2(-p.z, -p.z);
        ci[r] = a.s.cls_of_atom[st.sc_start + k];
        e_lj[r] = 0.0; e_c[r] = 0.0; clash[r] = false;
    }
    // warp-uniform amount of work: rows of the tile this warp has atoms in
    int rows = 0;
#pragma unroll
    for (int r = 0; r < SCORE_R; ++r) rows += (r * SCORE_THREADS + (tid & ~31) < n_local) ? 1 : 0;

    // environment chunk of this block (active order = class-major prefix order)
    const int n_split = SCREEN ? 1 : a.n_split;
    const int per = (n_active + n_split - 1) / n_split;
    const int lo = SCREEN ? 0 : blockIdx.z * per;
    const int hi = min(n_active, lo + per);
    const float4 *env = a.env + (long long)set * a.s.n_atoms;
    const int *scr_slot = SCREEN ? a.s.scr_slot + a.s.scr_off[a.step] : nullptr;

    // Layout of a tile: per class the atoms of [tile_lo, tile_hi) that belong to it, padded to an even count.  Warp 0
    // scans the padded lengths (s_seg) and the number of 32-atom staging chunks (s_chunk) of the classes, two tiles ahead
    // of the one being scored (three buffers).
    auto scan_tile = [&](int buf, int t_lo, int t_hi) {
        int carry_seg = 0, carry_chunk = 0;
        for (int c0 = 0; c0 < n_cls; c0 += 32) {
            const int c = c0 + lane;
            int len = 0;
            if (c < n_cls) len = max(0, min(s_prefix[c + 1], t_hi) - max(s_prefix[c], t_lo));
            const int even = (len + 1) & ~1, chunks = (even + 31) >> 5;
            int xs = even, xc = chunks;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int ys = __shfl_up_sync(0xffffffffu, xs, off), yc = __shfl_up_sync(0xffffffffu, xc, off);
                if (lane >= off) { xs += ys; xc += yc; }
            }
            if (c < n_cls) { s_seg3[buf][c] = carry_seg + xs - even; s_chunk3[buf][c] = carry_chunk + xc - chunks; }
            carry_seg += __shfl_sync(0xffffffffu, xs, 31);
            carry_chunk += __shfl_sync(0xffffffffu, xc, 31);
        }
        if (lane == 0) { s_seg3[buf][n_cls] = carry_seg; s_chunk3[buf][n_cls] = carry_chunk; }
    };
    // Staging of one tile, asynchronous (LDGSTS: global -> shared without registers, four 4-byte copies per atom into the
    // pair-interleaved layout): a warp takes whole chunks, so the class of a chunk is found once per 32 atoms.  The copies
    // of tile t+1 are issued before tile t is scored, so their latency hides under the pair loop.
    auto stage_tile = [&](int t, int t_lo, int t_hi) {
        const int *seg = s_seg3[t % 3], *chunk = s_chunk3[t % 3];
        float *flat = reinterpret_cast<float *>(s_env2[t & 1]);
        const int n_chunks = chunk[n_cls];
        int c = 0;
        for (int q = warp; q < n_chunks; q += SCORE_THREADS / 32) {
            while (chunk[c + 1] <= q) ++c;
            const int first = max(s_prefix[c], t_lo);                 // position of the class's first atom of this tile
            const int len = min(s_prefix[c + 1], t_hi) - first;
            const int e = ((q - chunk[c]) << 5) + lane;
            if (e < ((len + 1) & ~1)) {
                const int P = seg[c] + e, pp = P >> 1, hh = P & 1;
                float *dst = flat + 8 * pp + hh;
                if (e < len) {
                    const float *src = reinterpret_cast<const float *>(env + (SCREEN ? scr_slot[first + e] : a.s.cls_start[c] + (first - s_prefix[c]) + e));
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(src));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 8), "l"(src + 1));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 16), "l"(src + 2));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 24), "l"(src + 3));
                } else {                                               // pad slot of an odd segment: a parked dummy
                    dst[0] = 1.0e15f; dst[2] = 1.0e15f; dst[4] = 1.0e15f; dst[6] = 0.f;
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    const int n_tiles = (hi - lo + TILE_E - 1) / TILE_E;
    if (warp == 0) {
        if (n_tiles > 0) scan_tile(0, lo, min(lo + TILE_E, hi));
        if (n_tiles > 1) scan_tile(1, lo + TILE_E, min(lo + 2 * TILE_E, hi));
    }
    __syncthreads();
    if (n_tiles > 0) stage_tile(0, lo, min(lo + TILE_E, hi));
    for (int t = 0; t < n_tiles; ++t) {
        const int tile_lo = lo + t * TILE_E;
        const int tile_n = min(TILE_E, hi - tile_lo), tile_hi = tile_lo + tile_n;
        const int *s_seg = s_seg3[t % 3], *s_chunk = s_chunk3[t % 3];
        const float4 *s_env = s_env2[t & 1];
        float *s_flat = reinterpret_cast<float *>(s_env2[t & 1]);
        if (t + 1 < n_tiles) {                                            // prefetch the next tile into the other buffer
            stage_tile(t + 1, tile_hi, min(tile_hi + TILE_E, hi));
            asm volatile("cp.async.wait_group 1;" ::: "memory");          // this thread's copies of tile t have landed
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        // the special environment atoms are scored by the special-pair kernel: the thread that staged one parks it
        if (!SCREEN)
            for (int j = 0; j < st.n_sp_env; ++j) {
                const int ae = s_sp[j], cj = s_sp_c[j];
                if (ae < tile_lo || ae >= tile_hi) continue;
                const int e = ae - max(s_prefix[cj], tile_lo);
                if ((s_chunk[cj] + (e >> 5)) % (SCORE_THREADS / 32) == warp && (e & 31) == lane) {
                    const int P = s_seg[cj] + e, pp = P >> 1, hh = P & 1;
                    s_flat[8 * pp + hh] = 1.0e15f; s_flat[8 * pp + 2 + hh] = 1.0e15f; s_flat[8 * pp + 4 + hh] = 1.0e15f; s_flat[8 * pp + 6 + hh] = 0.f;
                }
            }
        __syncthreads();          // tile t staged by everybody
        if (warp == 0 && t + 2 < n_tiles) scan_tile((t + 2) % 3, tile_lo + 2 * TILE_E, min(tile_lo + 3 * TILE_E, hi));
        int c_first, c_last;
        { int c_lo = 0, c_hi = n_cls; while (c_hi - c_lo > 1) { int m = (c_lo + c_hi) >> 1; if (s_prefix[m] <= tile_lo) c_lo = m; else c_hi = m; } c_first = c_lo; }
        { int c_lo = 0, c_hi = n_cls; const int last = tile_lo + tile_n - 1;
          while (c_hi - c_lo > 1) { int m = (c_lo + c_hi) >> 1; if (s_prefix[m] <= last) c_lo = m; else c_hi = m; } c_last = c_lo; }
        if (SCREEN) {
            if (rows == SCORE_R) clash_tile<SCORE_R>(a, s_env, s_seg, c_first, c_last, nx, ny, nz, ci, clash);
            else if (SCORE_R > 2 && rows >= 2) clash_tile<(SCORE_R > 2 ? SCORE_R - 1 : 1)>(a, s_env, s_seg, c_first, c_last, nx, ny, nz, ci, clash);
            else if (rows >= 1) clash_tile<1>(a, s_env, s_seg, c_first, c_last, nx, ny, nz, ci, clash);
        } else {
            if (rows == SCORE_R) score_tile<SCORE_R>(a, s_env, s_seg, c_first, c_last, nx, ny, nz, ci, e_lj, e_c, clash);
            else if (SCORE_R > 2 && rows >= 2) score_tile<(SCORE_R > 2 ? SCORE_R - 1 : 1)>(a, s_env, s_seg, c_first, c_last, nx, ny, nz, ci, e_lj, e_c, clash);
            else if (rows >= 1) score_tile<1>(a, s_env, s_seg, c_first, c_last, nx, ny, nz, ci, e_lj, e_c, clash);
        }
        if (t + 2 < n_tiles) __syncthreads();     // tile t's buffer is overwritten by the prefetch of tile t+2; its scan is visible
    }

    // per-atom totals -> per-trial partial sums (fixed order)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SCORE_R; ++r) {
        const int la = r * SCORE_THREADS + tid;
        if (la < n_local) {
            if (!SCREEN) {
                const double qi = a.s.charge[st.sc_start + katom[r]];
                s_e[la] = e_lj[r] + COULOMB_PREF * qi * e_c[r];
            }
            s_clash[la] = clash[r] ? 1 : 0;
        }
    }
    __syncthreads();
    if (SCREEN) {
        int kept = 0;
        for (int t = tid; t < nt; t += SCORE_THREADS) {
            int cl = 0;
            for (int k = 0; k < n_per; ++k) cl |= s_clash[t * n_per + k];
            if (a.sp_partial[(long long)set * a.n_trials + t0 + t].y != 0.0) cl = 1;     // clash among the special pairs
            a.screen[(long long)set * a.n_trials + t0 + t] = (unsigned char)cl;
            kept += cl ? 0 : 1;
        }
        kept = __reduce_add_sync(0xffffffffu, kept);
        if (lane == 0 && kept) atomicAdd(a.n_keep + set, kept);
        return;
    }
    const int stride = a.n_trials + (a.dedup ? 1 : 0);
    for (int t = tid; t < nt; t += SCORE_THREADS) {
        double e = 0.0; int cl = 0;
        for (int k = 0; k < n_per; ++k) { e += s_e[t * n_per + k]; cl |= s_clash[t * n_per + k]; }
        a.partial[((long long)set * stride + trial_of(t0 + t)) * a.n_split + blockIdx.z] = make_double2(e, cl ? 1.0 : 0.0);
    }
    if (n_fixatoms > 0 && tid == SCORE_THREADS - 1) {
        double e = 0.0; int cl = 0;
        for (int k = 0; k < n_fixatoms; ++k) { e += s_e[n_varatoms + k]; cl |= s_clash[n_varatoms + k]; }
        a.partial[((long long)set * stride + a.n_trials) * a.n_split + blockIdx.z] = make_double2(e, cl ? 1.0 : 0.0);
    }
}

// ------------------------------------------------------------------------------------------------
// K2s: special pairs in float64.  grid (ceil(n_trials/(8*SP_TPW)), n_sets), block 256: one warp per SP_TPW trials
// ------------------------------------------------------------------------------------------------
struct SpecialArgs {
    SysDev s;
    int step;
    int n_trials;
    const float4 *env;
    const float4 *trial;
    long long trial_set_stride;
    double2 *sp_partial;               // [set][n_trials]
    const unsigned char *alive;
};

__device__ __forceinline__ double pair_energy_f64(float d2f, double A, double B, double Q, double thr, bool &clash) {
    // d = sqrt(float64(d2)); clash = d < thr; e = A/d^12 - B/d^6 + Q/d   (libenergy.py:683-688, 720, 798)
    const double d2 = (double)d2f;
    const double d = sqrt(d2);
    clash = d < thr;
    const double d6 = d2 * d2 * d2;
    double e = 0.0;
    if (A != 0.0 || B != 0.0) e = A / (d6 * d6) - B / d6;
    if (Q != 0.0) e += Q / d;
    return e;
}

// Special-pair energy with no float64 sqrt/div: 1/d from the float32 rsqrt refined by two Newton steps in
// float64 (relative error ~1e-15), then A/d^12 - B/d^6 + Q/d as products.  The clash decision is the exact
// float32 comparison d2_ref < thr2f (thr2f precomputed so that it equals sqrt(float64(d2_ref)) < thr).
__device__ __forceinline__ double pair_energy_fast(float d2f, double A, double B, double Q) {
    const double d2 = (double)d2f;
    double y = (double)rsqrtf(d2f);
    const double h = 0.5 * d2;
    y = y * (1.5 - h * y * y);
    y = y * (1.5 - h * y * y);
    const double w = y * y, w3 = w * w * w;
    return w3 * (A * w3 - B) + Q * y;
}

#define SP_TPW 4                        // trials per warp: the pair table entries a lane holds are reused for all of them
__global__ void __launch_bounds__(256) score_special_kernel(SpecialArgs a) {
    const int lane = threadIdx.x & 31;
    const int tb = (blockIdx.x * 8 + (threadIdx.x >> 5)) * SP_TPW;
    const int set = blockIdx.y;
    if (tb >= a.n_trials) return;
    if (a.alive && !a.alive[set]) return;
    const StepDev &st = a.s.steps[a.step];
    const float4 *trial0 = a.trial + (long long)set * a.trial_set_stride;
    const float4 *env = a.env + (long long)set * a.s.n_atoms;
    const int n_sc = st.n_sc;
    double e[SP_TPW]; bool any[SP_TPW];
#pragma unroll
    for (int i = 0; i < SP_TPW; ++i) { e[i] = 0.0; any[i] = false; }
    for (int p = lane; p < st.n_sp; p += 32) {
        const int q = st.sp_off + p;
        const int k = a.s.sp_k[q];
        const int pa = a.s.sp_partner[q];
        const double A = a.s.sp_A[q], B = a.s.sp_B[q], Q = a.s.sp_Q[q];
        const float thr2 = a.s.sp_thr2f[q];
        const bool live = (A != 0.0 || B != 0.0 || Q != 0.0);
        float4 bo_env = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pa >= 0) bo_env = env[st.sp_env_slot[pa]];          // environment partner: the same for every trial
#pragma unroll
        for (int i = 0; i < SP_TPW; ++i) {
            if (tb + i < a.n_trials) {
                const float4 *trial = trial0 + (long long)(tb + i) * n_sc;
                const float4 an = trial[k];
                const float4 bo = (pa >= 0) ? bo_env : trial[-pa - 1];
                const float d2 = dist2_ref(an.x, an.y, an.z, bo.x, bo.y, bo.z);
                any[i] |= d2 < thr2;
                if (live) e[i] += pair_energy_fast(d2, A, B, Q);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < SP_TPW; ++i) {
        double v = e[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        const bool cl = __any_sync(0xffffffffu, any[i]);
        if (lane == 0 && tb + i < a.n_trials) a.sp_partial[(long long)set * a.n_trials + tb + i] = make_double2(v, cl ? 1.0 : 0.0);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: clash mask + Boltzmann selection + append.  grid (n_sets), block 128, dyn smem 2*n_trials doubles
// ------------------------------------------------------------------------------------------------
struct SelectArgs {
    SysDev s;
    int step;
    int n_trials, n_split;
    const double2 *partial, *sp_partial;
    int dedup;
    const double *prob;                // [n_trials]
    const CallParams *cp;              // beta, seed, conformer offset (device memory)
    const double *uniform;             // explicit [set*u_stride + step] or null
    int u_stride;
    float4 *env;
    const float4 *trial;
    long long trial_set_stride;
    unsigned char *alive;
    double *e_acc;
    int *selected;                     // optional [set*n_res + step]
    double *trial_energy;              // optional dump [set*n_trials_total + trial_off + t]
    int *n_dead;                        // conformers that failed so far (read back now and then to re-size the grids)
    unsigned long long *pair_counter;   // [0] algorithmic pairs (the reference's P), [1] pairs actually evaluated
    long long pairs_per_set, pairs_computed_per_set;
    const unsigned char *screen;        // [set][n_trials] neighbour-screen flags (null = no screen): flagged trials are +inf and were not scored
    long long pairs_computed_per_kept;  // executed pairs of one trial the screen left (pairs_computed_per_set then holds the rest)
    int *n_keep;                        // [set] the screen's survivor counter, reset here for the next residue
    // coordinate-based term (hydrophobic PMF, libenergy.py:806-808): added to every trial that passed the clash filter
    const double *extra;                // [set][n_trials] or null
    float *gcoords;                     // [set][n_atoms][3] growth-order copy of the conformers (the term's kernels read it) or null
    const int *term_overflow;           // [set] the term exceeded a capacity for this conformer: it is given up (ok = 2)
};

// which (conformer, trial) pairs passed the clash filter: what the coordinate-based term is evaluated for
__global__ void __launch_bounds__(128) finite_mask_kernel(SelectArgs a, unsigned char *mask) {
    const int set = blockIdx.y;
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= a.n_trials) return;
    unsigned char ok = 0;
    if (!(a.alive && !a.alive[set]) && !(a.screen && a.screen[(long long)set * a.n_trials + t])) {
        const int stride = a.n_trials + (a.dedup ? 1 : 0);
        bool cl = a.sp_partial[(long long)set * a.n_trials + t].y != 0.0;
        for (int zz = 0; zz < a.n_split && !cl; ++zz) {
            cl |= a.partial[((long long)set * stride + t) * a.n_split + zz].y != 0.0;
            if (a.dedup) cl |= a.partial[((long long)set * stride + a.n_trials) * a.n_split + zz].y != 0.0;
        }
        ok = cl ? 0 : 1;
    }
    mask[(long long)set * a.n_trials + t] = ok;
}

__global__ void __launch_bounds__(128) select_kernel(SelectArgs a) {
    extern __shared__ double s_dyn[];
    double *s_E = s_dyn, *s_w = s_dyn + a.n_trials;
    __shared__ double s_red[4];
    __shared__ int s_kept[4];
    __shared__ int s_sel;
    const int set = blockIdx.x, tid = threadIdx.x;
    if (a.alive && !a.alive[set]) {
        if (a.selected && tid == 0) a.selected[(long long)set * a.s.n_res + a.step] = -1;
        return;
    }
    const StepDev &st = a.s.steps[a.step];
    if (a.n_keep && tid == 0) a.n_keep[set] = 0;
    double mn = CUDART_INF;
    const int stride = a.n_trials + (a.dedup ? 1 : 0);
    double e_fix = 0.0; bool cl_fix = false;
    if (a.dedup)
        for (int zz = 0; zz < a.n_split; ++zz) {
            const double2 p = a.partial[((long long)set * stride + a.n_trials) * a.n_split + zz];
            e_fix += p.x; cl_fix |= (p.y != 0.0);
        }
    int kept = 0;
    for (int t = tid; t < a.n_trials; t += 128) {
        double e = CUDART_INF;
        if (!(a.screen && a.screen[(long long)set * a.n_trials + t])) {
            const double2 sp = a.sp_partial[(long long)set * a.n_trials + t];
            e = sp.x; bool cl = (sp.y != 0.0) || cl_fix;
            for (int zz = 0; zz < a.n_split; ++zz) {
                const double2 p = a.partial[((long long)set * stride + t) * a.n_split + zz];
                e += p.x; cl |= (p.y != 0.0);
            }
            e += e_fix;
            if (cl) e = CUDART_INF;                           // energies[clash] = inf (libenergy.py:801)
            else if (a.extra) e += a.extra[(long long)set * a.n_trials + t];    // efuncs_coords (libenergy.py:806-808)
            ++kept;
        }
        s_E[t] = e;
        if (a.trial_energy) a.trial_energy[(long long)set * a.s.n_trials_total + st.trial_off + t] = e;
        mn = fmin(mn, e);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        kept += __shfl_xor_sync(0xffffffffu, kept, off);
    }
    if ((tid & 31) == 0) { s_red[tid >> 5] = mn; s_kept[tid >> 5] = kept; }
    __syncthreads();
    mn = fmin(fmin(s_red[0], s_red[1]), fmin(s_red[2], s_red[3]));
    if (tid == 0 && a.pair_counter) {
        kept = s_kept[0] + s_kept[1] + s_kept[2] + s_kept[3];
        atomicAdd(a.pair_counter, (unsigned long long)a.pairs_per_set);
        atomicAdd(a.pair_counter + 1, (unsigned long long)(a.pairs_computed_per_set + kept * a.pairs_computed_per_kept));
    }
    if (isinf(mn) || isnan(mn) || (a.term_overflow && a.term_overflow[set])) {   // every trial clashes: growth fails (side_chain_builder.py:187)
        if (tid == 0) {
            a.alive[set] = 0;
            if (a.n_dead) atomicAdd(a.n_dead, 1);
            if (a.selected) a.selected[(long long)set * a.s.n_res + a.step] = -1;
        }
        return;
    }
    for (int t = tid; t < a.n_trials; t += 128)
        s_w[t] = a.prob[t] * exp(-a.cp->beta * (s_E[t] - mn));    // side_chain_builder.py:191-192
    __syncthreads();
    if (tid == 0) {
        double u;
        if (a.uniform) u = a.uniform[(long long)set * a.u_stride + a.step];
        else {
            const long long cid = a.cp->conf_id0 + set;
            const unsigned long long seed = a.cp->seed;
            uint4 r = philox4x32(make_uint4((unsigned)cid, (unsigned)(cid >> 32), (unsigned)a.step, 0xFFFFFFFFu),
                                 make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
            u = u01_halfopen(r.x, r.y);
        }
        // sequential float64 running sums, first index whose sum exceeds u*total (side_chain_builder.py:35-40)
        double total = 0.0;
        for (int t = 0; t < a.n_trials; ++t) total += s_w[t];
        const double pointer = u * total;
        int sel = a.n_trials - 1;
        double run = 0.0;
        for (int t = 0; t < a.n_trials; ++t) { run += s_w[t]; if (run > pointer) { sel = t; break; } }
        s_sel = sel;
        a.e_acc[set] += (s_E[sel] - mn) + mn;                 // side_chain_builder.py:197
        if (a.selected) a.selected[(long long)set * a.s.n_res + a.step] = sel;
    }
    __syncthreads();
    const int sel = s_sel;
    if (tid < st.n_sc) {
        float4 p = a.trial[(long long)set * a.trial_set_stride + (long long)sel * st.n_sc + tid];
        const int atom = st.sc_start + tid;
        p.w = (float)a.s.charge[atom];
        a.env[(long long)set * a.s.n_atoms + a.s.slot_of_atom[atom]] = p;
        if (a.gcoords) {
            float *g = a.gcoords + ((long long)set * a.s.n_atoms + atom) * 3;
            g[0] = p.x; g[1] = p.y; g[2] = p.z;
        }
    }
}

// stand-alone selection on caller-provided energies (C-ABI mcsce_select)
__global__ void __launch_bounds__(128) select_plain_kernel(int n_trials, const double *energy, const double *prob, double beta,
                                                           const double *uniform, int *selected, double *e_sel) {
    extern __shared__ double s_w[];
    __shared__ double s_red[4];
    const int set = blockIdx.x, tid = threadIdx.x;
    const double *E = energy + (long long)set * n_trials;
    double mn = CUDART_INF;
    for (int t = tid; t < n_trials; t += 128) mn = fmin(mn, E[t]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    if ((tid & 31) == 0) s_red[tid >> 5] = mn;
    __syncthreads();
    mn = fmin(fmin(s_red[0], s_red[1]), fmin(s_red[2], s_red[3]));
    if (isinf(mn) || isnan(mn)) { if (tid == 0) { selected[set] = -1; e_sel[set] = CUDART_INF; } return; }
    for (int t = tid; t < n_trials; t += 128) s_w[t] = prob[t] * exp(-beta * (E[t] - mn));
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int t = 0; t < n_trials; ++t) total += s_w[t];
        const double pointer = uniform[set] * total;
        int sel = n_trials - 1; double run = 0.0;
        for (int t = 0; t < n_trials; ++t) { run += s_w[t]; if (run > pointer) { sel = t; break; } }
        selected[set] = sel; e_sel[set] = (E[sel] - mn) + mn;
    }
}

// ------------------------------------------------------------------------------------------------
// environment init, input-atom self energy, result gather
// ------------------------------------------------------------------------------------------------
__global__ void init_env_kernel(SysDev s, int n_sets, float4 *env) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_sets * s.n_atoms) return;
    const int atom = (int)(i % s.n_atoms);
    const int set = (int)(i / s.n_atoms);
    float4 v = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.f);
    const float q = (float)s.charge[atom];
    if (atom < s.n_input) v = make_float4(s.input_xyz[atom * 3], s.input_xyz[atom * 3 + 1], s.input_xyz[atom * 3 + 2], q);
    else {
        // GLY/ALA side chains are conformer independent: placed once (side_chain_builder.py:164-168)
        const float rx = s.rigid_xyz[atom * 3];
        if (!isnan(rx)) v = make_float4(rx, s.rigid_xyz[atom * 3 + 1], s.rigid_xyz[atom * 3 + 2], q);
    }
    env[(long long)set * s.n_atoms + s.slot_of_atom[atom]] = v;
}

__global__ void gather_kernel(SysDev s, int n_sets, const float4 *env, float *coords) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_sets * s.n_atoms) return;
    const int atom = (int)(i % s.n_atoms);
    const int set = (int)(i / s.n_atoms);
    const float4 v = env[(long long)set * s.n_atoms + s.slot_of_atom[atom]];
    coords[i * 3] = v.x; coords[i * 3 + 1] = v.y; coords[i * 3 + 2] = v.z;
}

// generic input-input pairs, one thread per row i, float64; same/adjacent residue numbers are in bb list
__global__ void input_energy_rows_kernel(SysDev s, double2 *row_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n_input) return;
    const float ax = s.input_xyz[i * 3], ay = s.input_xyz[i * 3 + 1], az = s.input_xyz[i * 3 + 2];
    const int ci = s.cls_of_atom[i], ri = s.resnum[i];
    const double qi = s.charge[i];
    double e = 0.0; bool any = false;
    for (int j = i + 1; j < s.n_input; ++j) {
        const int dr = s.resnum[j] - ri;
        if (dr >= -1 && dr <= 1) continue;
        const int cj = s.cls_of_atom[j];
        const float d2 = dist2_ref(ax, ay, az, s.input_xyz[j * 3], s.input_xyz[j * 3 + 1], s.input_xyz[j * 3 + 2]);
        const double s2 = s.cls_s2d[ci * s.n_cls + cj], e4 = s.cls_e4[ci * s.n_cls + cj];
        const double s6 = s2 * s2 * s2;
        bool cl;
        e += pair_energy_f64(d2, e4 * s6 * s6, e4 * s6, COULOMB_PREF * qi * s.charge[j], s.cls_thrd[ci * s.n_cls + cj], cl);
        any |= cl;
    }
    row_out[i] = make_double2(e, any ? 1.0 : 0.0);
}

__global__ void input_energy_pairs_kernel(SysDev s, double2 *pair_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n_bb_pairs) return;
    const int i = s.bb_i[p], j = s.bb_j[p];
    const double *cf = s.bb_coef + (long long)p * 4;
    const float d2 = dist2_ref(s.input_xyz[i * 3], s.input_xyz[i * 3 + 1], s.input_xyz[i * 3 + 2],
                               s.input_xyz[j * 3], s.input_xyz[j * 3 + 1], s.input_xyz[j * 3 + 2]);
    bool cl;
    const double e = pair_energy_f64(d2, cf[0], cf[1], cf[2], cf[3], cl);
    pair_out[p] = make_double2(e, cl ? 1.0 : 0.0);
}

__global__ void input_energy_reduce_kernel(int n_rows, const double2 *rows, int n_pairs, const double2 *pairs, double *out) {
    // single block, fixed-order tree: deterministic
    __shared__ double s_sum[256];
    __shared__ int s_cl[256];
    double e = 0.0; int cl = 0;
    for (int i = threadIdx.x; i < n_rows; i += 256) { e += rows[i].x; cl |= rows[i].y != 0.0; }
    for (int i = threadIdx.x; i < n_pairs; i += 256) { e += pairs[i].x; cl |= pairs[i].y != 0.0; }
    s_sum[threadIdx.x] = e; s_cl[threadIdx.x] = cl;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s_sum[threadIdx.x] += s_sum[threadIdx.x + w]; s_cl[threadIdx.x] |= s_cl[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s_cl[0] ? CUDART_INF : s_sum[0];
}

__global__ void init_state_kernel(int n_sets, unsigned char *alive, double *e_acc, const double *e_input, const double *e_term,
                                  int *term_overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_sets) {
        alive[i] = 1; e_acc[i] = e_term ? e_input[0] + e_term[0] : e_input[0];
        if (term_overflow) term_overflow[i] = 0;
    }
}

__global__ void finish_kernel(int n_sets, const unsigned char *alive, const double *e_acc, double *energy, unsigned char *ok,
                              const int *term_overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_sets) {
        ok[i] = (term_overflow && term_overflow[i]) ? 2 : alive[i];
        energy[i] = alive[i] ? e_acc[i] : CUDART_NAN;
    }
}

// ------------------------------------------------------------------------------------------------
// K0: the whole growth in one kernel, a thread-block cluster per conformer (launch-bound runs)
// ------------------------------------------------------------------------------------------------
#include "grow_cluster.cuh"
#include "score_conf.cuh"

// ------------------------------------------------------------------------------------------------
// host side: handle, table upload, launch logic
// ------------------------------------------------------------------------------------------------
#ifndef MAX_LANES
#define MAX_LANES 4
#endif
#ifndef LANE_LIMIT
#define LANE_LIMIT 4                    // conformer groups in flight at most (round 2, interleaved enqueue: 2 / 3 / 4 groups
                                        // 77.7 / 76.0 / 74.8 ms at 256 conformers of config 3, 539.9 / 537.2 / 536.0 ms at 2048)
#endif

struct mcsce_handle_s {
    int device = 0;
    int sm_count = 148;
    std::vector<void *> sys_allocs;
    char *bb_slab = nullptr;              // backbone-dependent tables: one grow-only device allocation, refilled by set_backbone
    char *bb_stage = nullptr;             // its pinned host image: one copy per backbone
    size_t bb_cap = 0;
    const double2 *d_bb_rows = nullptr, *d_bb_pairs = nullptr;   // scratch of the input-atom energy (inside the slab)
    cudaEvent_t bb_ready = nullptr;       // recorded when set_backbone's copies and kernels are enqueued
    SysDev d{};
    bool has_sys = false, has_bb = false;
    bool warned_screen_limit = false;
    TileRule rule{SCORE_R, 1, 0.4f};       // packed 2-row tiles; tuning knobs: MCSCE_B200_TILE_ROWS / _TILE_PACKED / _TILE_OVERHEAD
    // host mirrors needed for launch geometry
    std::vector<StepDev> steps;
    std::vector<TmplDev> tmpl_host;
    std::vector<int> cls_active, slot_of_atom, cls_of_atom, cls_start;
    std::vector<double> cls_thr;
    std::vector<int> resnum;
    std::vector<int> step_n_sp_env;
    std::vector<int> scr_len;            // neighbour-screen list length per residue (0 = residue not screened)
    int max_trials = 0, max_trial_atoms = 0;
    StepDev *d_steps = nullptr;
    unsigned long long *d_pairs = nullptr;
    int *h_n_dead = nullptr;                           // pinned host copy of a lane's dead-conformer counter
    // cached CUDA graph of one growth (small, launch-bound runs)
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<long long> graph_key;
    long long graph_launches = 0;
    long long table_version = 0;
    long long launches = 0;
    // A lane = what one group of conformers in flight needs: a main stream (lane 0: the caller's), a second stream for
    // the placement of the next residue and the special-pair kernel, their events, the call's parameter block.  Big calls
    // run two groups side by side so that the launch tails and the small kernels of one hide under the other's scoring pass.
    std::vector<float> input_xyz_host;               // backbone of the last set_backbone (near lists of the hydrophobic PMF)
    std::vector<float> hp_reach;                     // per residue: upper bound of |side-chain atom - CA| (reach_bounds)
    struct Lane {
        // buffers of the coordinate-based term (grow-only): growth-order coordinates, per-trial values, mask, overflow flags
        float *hp_gcoords = nullptr; double *hp_v = nullptr, *hp_v_init = nullptr; unsigned char *hp_mask = nullptr; int *hp_ovf = nullptr;
        size_t hp_cap_conf = 0, hp_cap_atoms = 0, hp_cap_trials = 0;
        cudaStream_t main = nullptr, aux = nullptr;
        std::vector<cudaEvent_t> evP, evQ, evS;
        cudaEvent_t ev0 = nullptr, done = nullptr;
        CallParams *d_call = nullptr;
        int *d_n_dead = nullptr;
    } lane[MAX_LANES];
    double *d_input_energy = nullptr;         // energy of the input atoms alone: a constant of the backbone, computed in set_backbone
    // persistent cluster growth (grow_cluster.cuh): layout of the shared-memory environment image, cached launch plan
    const int *d_pstart = nullptr;            // [n_cls+1] even-padded class starts
    const int *d_aoff = nullptr;              // [n_res+1] first trial atom of every residue (all trials of a conformer in one block)
    long long trial_atoms_total = 0;
    int env_pairs = 0;
    long long gc_plan_key[2] = {-1, -1};      // (table_version, n_conf) the cached plan was made for
    int gc_S = 0, gc_threads = 0, gc_grid = 0;
    size_t gc_smem = 0;
    long long persistent_launches = 0;
    bool gc_np_set = false;                   // clusters of 16 CTAs allowed on this handle's device
    int k2conf = -1;                          // scoring pass with one CTA per conformer (score_conf.cuh): -1 = not read yet
    size_t k2conf_smem = 0;
    const CallParams *cur_call = nullptr;     // parameter block of the lane being enqueued (read by launch_place)
    // optional per-kernel-class device timing (CUDA events on the launch stream)
    bool profiling = false;
    bool capturing = false;
    std::vector<cudaEvent_t> ev[5];     // 0 place, 1 score_generic, 2 score_special, 3 select, 4 hpmf: begin/end pairs
};

struct ProfScope {
    mcsce_handle h; int cls; cudaStream_t st;
    ProfScope(mcsce_handle h_, int cls_, cudaStream_t st_) : h(h_), cls(cls_), st(st_) {
        if (h->profiling) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); h->ev[cls].push_back(e); }
    }
    ~ProfScope() {
        if (h->profiling) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); h->ev[cls].push_back(e); }
    }
};

template <typename T>
static int upload(mcsce_handle h, std::vector<void *> &pool, const T *src, size_t n, const T **dst) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    pool.push_back(p);
    if (n) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return 0;
}
static void free_pool(std::vector<void *> &pool) { for (void *p : pool) cudaFree(p); pool.clear(); }

// float32 t2 such that for every float32 d2: (d2 < t2) <=> (sqrt((double)d2) < thr)
static float exact_thr2(double thr) {
    if (!(thr > 0.0)) return 0.0f;
    auto pred = [&](float x) { return sqrt((double)x) < thr; };
    float x = (float)(thr * thr);
    while (pred(x)) x = nextafterf(x, INFINITY);
    while (x > 0.f && !pred(nextafterf(x, 0.f))) x = nextafterf(x, 0.f);
    return x;
}

// Dynamic shared memory of the selection kernels: 2 (select_kernel) or 1 (select_plain_kernel) doubles per trial.  Above the
// 48 KB default the kernel needs an explicit opt-in; above the 227 KB a CTA can have the trial set is refused with a message
// (the reference's largest sets: 81 x 9 = 729 rows for ARG/LYS, 288 for M3L/KCX - far below either limit).
#define MAX_DYN_SMEM (227 * 1024 - 2048)
template <typename K>
static int ensure_dyn_smem(K kernel, size_t bytes, const char *what) {
    if (bytes <= 48 * 1024) return 0;
    if (bytes > MAX_DYN_SMEM)
        return fail(std::string(what) + ": trial set too large for the selection kernel's shared memory (" +
                    std::to_string(bytes / 1024) + " KB needed, " + std::to_string(MAX_DYN_SMEM / 1024) + " KB available)");
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

extern "C" const char *mcsce_last_error(void) { return g_err.c_str(); }
extern "C" int mcsce_version(void) { return 1; }

extern "C" int mcsce_create(mcsce_handle *out, int device) {
    if (!out) return fail("null handle pointer");
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail("no such CUDA device");
    CK(cudaSetDevice(device));
    mcsce_handle h = new mcsce_handle_s();
    h->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    CK(cudaMalloc((void **)&h->d_pairs, 2 * sizeof(unsigned long long)));
    CK(cudaMemset(h->d_pairs, 0, 2 * sizeof(unsigned long long)));
    for (auto &ln : h->lane) {
        CK(cudaMalloc((void **)&ln.d_n_dead, sizeof(int)));
        CK(cudaMalloc((void **)&ln.d_call, sizeof(CallParams)));
        CK(cudaMemset(ln.d_call, 0, sizeof(CallParams)));
    }
    h->cur_call = h->lane[0].d_call;
    if (const char *e = getenv("MCSCE_B200_TILE_ROWS")) h->rule.max_rows = std::max(1, std::min(SCORE_R, atoi(e)));
    if (const char *e = getenv("MCSCE_B200_TILE_PACKED")) h->rule.packed = atoi(e) ? 1 : 0;
    if (const char *e = getenv("MCSCE_B200_TILE_OVERHEAD")) h->rule.overhead = (float)atof(e);
    CK(cudaMalloc((void **)&h->d_input_energy, sizeof(double)));
    CK(cudaMallocHost((void **)&h->h_n_dead, sizeof(int) * MAX_LANES));
    CK(cudaEventCreateWithFlags(&h->bb_ready, cudaEventDisableTiming));
    *out = h;
    return 0;
}

extern "C" int mcsce_destroy(mcsce_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    free_pool(h->sys_allocs);
    if (h->bb_slab) cudaFree(h->bb_slab);
    if (h->bb_stage) cudaFreeHost(h->bb_stage);
    if (h->bb_ready) cudaEventDestroy(h->bb_ready);
    for (auto &ln : h->lane) {
        for (auto e : ln.evP) cudaEventDestroy(e);
        for (auto e : ln.evQ) cudaEventDestroy(e);
        for (auto e : ln.evS) cudaEventDestroy(e);
        if (ln.ev0) cudaEventDestroy(ln.ev0);
        if (ln.done) cudaEventDestroy(ln.done);
        if (ln.aux) cudaStreamDestroy(ln.aux);
        if (ln.main) cudaStreamDestroy(ln.main);
        if (ln.d_n_dead) cudaFree(ln.d_n_dead);
        if (ln.d_call) cudaFree(ln.d_call);
        if (ln.hp_gcoords) cudaFree(ln.hp_gcoords);
        if (ln.hp_v) cudaFree(ln.hp_v);
        if (ln.hp_v_init) cudaFree(ln.hp_v_init);
        if (ln.hp_mask) cudaFree(ln.hp_mask);
        if (ln.hp_ovf) cudaFree(ln.hp_ovf);
    }
    if (h->d_input_energy) cudaFree(h->d_input_energy);
    if (h->d_pairs) cudaFree(h->d_pairs);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->h_n_dead) cudaFreeHost(h->h_n_dead);
    delete h;
    return 0;
}

extern "C" int mcsce_set_system(mcsce_handle h, const mcsce_system_desc *s) {
    if (!h || !s) return fail("null argument");
    CK(cudaSetDevice(h->device));
    if (s->n_cls > MAX_CLS) return fail("too many atom classes");
    free_pool(h->sys_allocs);
    h->has_sys = h->has_bb = false;
    h->table_version++;
    SysDev &d = h->d;
    d = SysDev{};
    d.n_atoms = s->n_atoms; d.n_input = s->n_input; d.n_res = s->n_res; d.n_cls = s->n_cls; d.n_types = s->n_types;
    auto &P = h->sys_allocs;
    const size_t nc2 = (size_t)s->n_cls * s->n_cls;
    if (upload(h, P, s->slot_of_atom, s->n_atoms, &d.slot_of_atom)) return -1;
    if (upload(h, P, s->cls_of_atom, s->n_atoms, &d.cls_of_atom)) return -1;
    if (upload(h, P, s->charge, s->n_atoms, &d.charge)) return -1;
    if (upload(h, P, s->resnum, s->n_atoms, &d.resnum)) return -1;
    if (upload(h, P, s->cls_start, s->n_cls + 1, &d.cls_start)) return -1;
    if (upload(h, P, s->cls_active, (size_t)(s->n_res + 1) * s->n_cls, &d.cls_active)) return -1;
    if (upload(h, P, s->cls_e4, nc2, &d.cls_e4)) return -1;
    if (upload(h, P, s->cls_s2, nc2, &d.cls_s2d)) return -1;
    if (upload(h, P, s->cls_thr, nc2, &d.cls_thrd)) return -1;
    std::vector<float> s2f(nc2), thr2f(nc2);
    for (size_t i = 0; i < nc2; ++i) { s2f[i] = (float)s->cls_s2[i]; thr2f[i] = exact_thr2(s->cls_thr[i]); }
    {
        std::vector<double> cA(nc2), cB(nc2);
        for (size_t i = 0; i < nc2; ++i) {
            const double s6 = s->cls_s2[i] * s->cls_s2[i] * s->cls_s2[i];
            cA[i] = s->cls_e4[i] * s6 * s6; cB[i] = s->cls_e4[i] * s6;
        }
        if (upload(h, P, cA.data(), nc2, &d.cls_A)) return -1;
        if (upload(h, P, cB.data(), nc2, &d.cls_B)) return -1;
    }
    if (upload(h, P, s2f.data(), nc2, &d.cls_s2f)) return -1;
    if (upload(h, P, thr2f.data(), nc2, &d.cls_thr2f)) return -1;
    h->cls_active.assign(s->cls_active, s->cls_active + (size_t)(s->n_res + 1) * s->n_cls);
    {
        std::vector<int> pre((size_t)(s->n_res + 1) * (s->n_cls + 1), 0);
        for (int i = 0; i <= s->n_res; ++i)
            for (int c = 0; c < s->n_cls; ++c)
                pre[(size_t)i * (s->n_cls + 1) + c + 1] = pre[(size_t)i * (s->n_cls + 1) + c] + s->cls_active[(size_t)i * s->n_cls + c];
        if (upload(h, P, pre.data(), pre.size(), &d.cls_prefix)) return -1;
    }
    {   // shared-memory environment image of the persistent growth: every class segment starts at an even position
        std::vector<int> ps(s->n_cls + 1, 0);
        for (int c = 0; c < s->n_cls; ++c) ps[c + 1] = ps[c] + ((s->cls_start[c + 1] - s->cls_start[c] + 1) & ~1);
        h->env_pairs = ps[s->n_cls] / 2;
        if (upload(h, P, ps.data(), ps.size(), &h->d_pstart)) return -1;
    }
    h->slot_of_atom.assign(s->slot_of_atom, s->slot_of_atom + s->n_atoms);
    h->cls_of_atom.assign(s->cls_of_atom, s->cls_of_atom + s->n_atoms);
    h->cls_start.assign(s->cls_start, s->cls_start + s->n_cls + 1);
    h->cls_thr.assign(s->cls_thr, s->cls_thr + nc2);
    h->resnum.assign(s->resnum, s->resnum + s->n_atoms);

    std::vector<TmplDev> tm(s->n_types);
    for (int t = 0; t < s->n_types; ++t) {
        TmplDev &o = tm[t];
        memset(&o, 0, sizeof(o));
        o.n_atoms = s->tmpl_n_atoms[t];
        if (o.n_atoms > MAX_TMPL) return fail("template too large");
        o.n_chi = s->tmpl_n_chi[t];
        for (int a = 0; a < MAX_TMPL; ++a) {
            o.sc_rank[a] = -1;
            for (int k = 0; k < 3; ++k) o.xyz[a][k] = s->tmpl_xyz[((size_t)t * MAX_TMPL + a) * 3 + k];
        }
        for (int k = 0; k < MAX_SC; ++k) {
            int pos = s->tmpl_sc_pos[(size_t)t * MAX_SC + k];
            if (pos >= 0) o.sc_rank[pos] = k;
        }
        for (int k = 0; k < MAX_ROT; ++k) {
            o.axis_from[k] = s->tmpl_axis[((size_t)t * MAX_ROT + k) * 2];
            o.axis_to[k] = s->tmpl_axis[((size_t)t * MAX_ROT + k) * 2 + 1];
            o.move[k] = s->tmpl_move[(size_t)t * MAX_ROT + k];
            o.ori[k] = s->tmpl_ori[(size_t)t * MAX_ROT + k];
        }
    }
    if (upload(h, P, tm.data(), tm.size(), &d.tmpl)) return -1;
    h->tmpl_host = tm;
    h->steps.assign(s->n_res, StepDev{});
    h->step_n_sp_env.assign(s->n_res, 0);
    for (int i = 0; i < s->n_res; ++i) {
        StepDev &o = h->steps[i];
        memset(&o, 0, sizeof(o));
        o.kind = s->step_kind[i]; o.type = s->step_type[i]; o.n_sc = s->step_n_sc[i]; o.sc_start = s->step_sc_start[i];
        o.n_sp_env = s->step_n_sp_env[i]; o.sp_off = s->step_sp_off[i]; o.n_sp = s->step_n_sp[i];
        if (o.n_sp_env > MAX_SP_ENV) return fail("too many special environment atoms for one residue");
        if (o.n_sc > MAX_SC) return fail("side chain too large");
        for (int k = 0; k < MAX_SP_ENV; ++k)
            o.sp_env_slot[k] = (k < o.n_sp_env) ? s->slot_of_atom[s->step_sp_env[(size_t)i * MAX_SP_ENV + k]] : -1;
    }
    if (upload(h, P, s->sp_k, s->n_sp_total, &d.sp_k)) return -1;
    if (upload(h, P, s->sp_partner, s->n_sp_total, &d.sp_partner)) return -1;
    if (upload(h, P, s->sp_coef, (size_t)s->n_sp_total * 4, &d.sp_coef)) return -1;
    for (int c = 0; c < 3; ++c) {
        std::vector<double> col((size_t)s->n_sp_total);
        for (int i = 0; i < s->n_sp_total; ++i) col[i] = s->sp_coef[(size_t)i * 4 + c];
        if (upload(h, P, col.data(), col.size(), c == 0 ? &d.sp_A : c == 1 ? &d.sp_B : &d.sp_Q)) return -1;
    }
    {
        std::vector<float> t2((size_t)s->n_sp_total);
        for (int i = 0; i < s->n_sp_total; ++i) t2[i] = exact_thr2(s->sp_coef[(size_t)i * 4 + 3]);
        if (upload(h, P, t2.data(), t2.size(), &d.sp_thr2f)) return -1;
    }
    d.n_bb_pairs = s->n_bb_pairs;
    if (upload(h, P, s->bb_i, s->n_bb_pairs, &d.bb_i)) return -1;
    if (upload(h, P, s->bb_j, s->n_bb_pairs, &d.bb_j)) return -1;
    if (upload(h, P, s->bb_coef, (size_t)s->n_bb_pairs * 4, &d.bb_coef)) return -1;
    d.steps = nullptr;                   // filled by set_backbone (the per-step records carry backbone-dependent fields)
    h->has_sys = true;
    return 0;
}

static int launch_place(mcsce_handle h, int step, int n_sets, int n_rows, const double *chis_in, long long set_stride,
                        long long row_off, double *chis_out, float4 *trial, long long trial_stride,
                        const unsigned char *alive, cudaStream_t st, const double *ori_in = nullptr);

// GLY/ALA side chains have no chi: the template is placed once per backbone and shared by every conformer
// (side_chain_builder.py:164-168).  One block per residue, lane = template atom; same arithmetic as K1's placement step.
__global__ void __launch_bounds__(32) place_rigid_kernel(SysDev s, float *rigid) {
    const StepDev &st = s.steps[blockIdx.x];
    if (st.kind != 1) return;
    const TmplDev &tm = s.tmpl[st.type];
    const int lane = threadIdx.x;
    if (lane >= tm.n_atoms) return;
    const int rank = tm.sc_rank[lane];
    if (rank < 0 || rank >= st.n_sc) return;
    double ox, oy, oz;
    quat_rotate(st.q1[0], st.q1[1], st.q1[2], st.q1[3], (double)tm.xyz[lane][0], (double)tm.xyz[lane][1], (double)tm.xyz[lane][2], ox, oy, oz);
    const float rx = __double2float_rn(ox), ry = __double2float_rn(oy), rz = __double2float_rn(oz);
    quat_rotate(st.q2[0], st.q2[1], st.q2[2], st.q2[3], (double)rx, (double)ry, (double)rz, ox, oy, oz);
    float *o = rigid + (size_t)(st.sc_start + rank) * 3;
    o[0] = __double2float_rn(__dadd_rn(ox, (double)st.ca[0]));
    o[1] = __double2float_rn(__dadd_rn(oy, (double)st.ca[1]));
    o[2] = __double2float_rn(__dadd_rn(oz, (double)st.ca[2]));
}

// Neighbour screen lists.  A trial that clashes with any atom is +inf no matter what else it touches, and most clashes
// are with atoms a few angstrom away, so each grown residue gets a short list of environment slots that can come near
// its side chain; K2 first tests every trial against that list (clash only) and scores just the survivors.  The list is
// a heuristic: the scoring pass still tests every pair it evaluates, so a miss costs time, never correctness.
//   input atoms: known position; side-chain atoms of earlier residues: within their template distance of their CA.
static int launch_input_energy(mcsce_handle h, double2 *rows, double2 *pairs, double *out, cudaStream_t st);

static int build_screen_lists(mcsce_handle h, const mcsce_backbone_desc *b, std::vector<int> &off, std::vector<int> &slots,
                              std::vector<int> &prefix) {
    SysDev &d = h->d;
    const int n_cls = d.n_cls, n_res = d.n_res;
    double scale = 1.0;
    if (const char *e = getenv("MCSCE_B200_SCREEN_SCALE")) scale = atof(e);      // tuning knob; 0 disables the screen
    off.assign(n_res + 1, 0); slots.clear(); prefix.assign((size_t)n_res * (n_cls + 1), 0);
    h->scr_len.assign(n_res, 0);
    // every atom: a centre and a radius that bound where it can be
    std::vector<float> cx((size_t)d.n_atoms * 3, 0.f), rad(d.n_atoms, 0.f);
    for (int a = 0; a < d.n_input; ++a) for (int k = 0; k < 3; ++k) cx[(size_t)a * 3 + k] = b->input_xyz[(size_t)a * 3 + k];
    std::vector<float> reach(n_res, 0.f);
    for (int j = 0; j < n_res; ++j) {
        const StepDev &sj = h->steps[j];
        if (sj.kind == 0) continue;
        const TmplDev &tm = h->tmpl_host[sj.type];
        for (int p = 0; p < tm.n_atoms; ++p) {
            const int k = tm.sc_rank[p];
            if (k < 0 || k >= sj.n_sc) continue;
            const float r = sqrtf(tm.xyz[p][0] * tm.xyz[p][0] + tm.xyz[p][1] * tm.xyz[p][1] + tm.xyz[p][2] * tm.xyz[p][2]);
            const int a = sj.sc_start + k;
            for (int q = 0; q < 3; ++q) cx[(size_t)a * 3 + q] = sj.ca[q];
            rad[a] = (sj.kind == 2) ? 1.05f * r + 0.2f : r + 0.05f;      // chi rotations change the distance to CA a little
            reach[j] = std::max(reach[j], rad[a]);
        }
    }
    // atom groups with a bounding sphere each (runs of input atoms with one residue number, one side chain per
    // residue), so that a residue only looks at the atoms of groups that can reach it
    struct Group { int a0, a1; float c[3], r; };
    std::vector<Group> groups;
    auto close_group = [&](int a0, int a1) {
        Group g{a0, a1, {0.f, 0.f, 0.f}, 0.f};
        for (int a = a0; a < a1; ++a) for (int k = 0; k < 3; ++k) g.c[k] += cx[(size_t)a * 3 + k] / (float)(a1 - a0);
        for (int a = a0; a < a1; ++a) {
            const float dx = cx[(size_t)a * 3] - g.c[0], dy = cx[(size_t)a * 3 + 1] - g.c[1], dz = cx[(size_t)a * 3 + 2] - g.c[2];
            g.r = std::max(g.r, sqrtf(dx * dx + dy * dy + dz * dz) + rad[a]);
        }
        groups.push_back(g);
    };
    for (int a = 0, a0 = 0; a <= d.n_input; ++a)
        if (a == d.n_input || (a > a0 && h->resnum[a] != h->resnum[a0]) || a - a0 >= 64) { if (a > a0) close_group(a0, a); a0 = a; }
    for (int j = 0; j < n_res; ++j)
        if (h->steps[j].kind != 0 && h->steps[j].n_sc > 0) close_group(h->steps[j].sc_start, h->steps[j].sc_start + h->steps[j].n_sc);
    double thr_max = 0.0;
    for (double v : h->cls_thr) thr_max = std::max(thr_max, v);
    std::vector<double> thr_i(n_cls);
    std::vector<std::vector<int>> bucket(n_cls);
    for (int i = 0; i < n_res && scale > 0.0; ++i) {
        off[i] = (int)slots.size();
        const StepDev &si = h->steps[i];
        if (si.kind == 2 && si.n_trials > SCREEN_WORDS * 32) {      // the screen's survivor mask holds SCREEN_WORDS x 32 trials
            if (!h->warned_screen_limit)
                fprintf(stderr, "[mcsce_b200] note: residue %d has %d trials (> %d): the neighbour screen is off for it, every trial "
                                "is scored against the whole environment (same results, more work)\n", i, si.n_trials, SCREEN_WORDS * 32);
            h->warned_screen_limit = true;
            continue;
        }
        if (si.kind != 2 || si.n_trials < 8) continue;
        int n_active = 0;
        for (int c = 0; c < n_cls; ++c) n_active += h->cls_active[(size_t)i * n_cls + c];
        if (n_active < 256) continue;
        std::fill(thr_i.begin(), thr_i.end(), 0.0);
        for (int k = 0; k < si.n_sc; ++k) {
            const int ck = h->cls_of_atom[si.sc_start + k];
            for (int c = 0; c < n_cls; ++c) thr_i[c] = std::max(thr_i[c], h->cls_thr[(size_t)ck * n_cls + c]);
        }
        for (auto &v : bucket) v.clear();
        int L = 0;
        for (const Group &g : groups) {
            const float gx = g.c[0] - si.ca[0], gy = g.c[1] - si.ca[1], gz = g.c[2] - si.ca[2];
            const double glim = scale * (reach[i] + thr_max) + g.r;
            if ((double)gx * gx + (double)gy * gy + (double)gz * gz >= glim * glim) continue;
            for (int a = g.a0; a < g.a1; ++a) {
                const int c = h->cls_of_atom[a], slot = h->slot_of_atom[a];
                if (slot - h->cls_start[c] >= h->cls_active[(size_t)i * n_cls + c]) continue;        // not placed yet
                if (!(thr_i[c] > 0.0)) continue;
                bool special = false;
                for (int q = 0; q < si.n_sp_env; ++q) special |= (si.sp_env_slot[q] == slot);
                if (special) continue;
                const float dx = cx[(size_t)a * 3] - si.ca[0], dy = cx[(size_t)a * 3 + 1] - si.ca[1], dz = cx[(size_t)a * 3 + 2] - si.ca[2];
                const double lim = scale * (reach[i] + rad[a] + thr_i[c]);
                if ((double)dx * dx + (double)dy * dy + (double)dz * dz < lim * lim) { bucket[c].push_back(slot); ++L; }
            }
        }
        if (L == 0 || (long long)L * 5 > (long long)n_active * 2) continue;     // no gain when the list is most of the environment
        for (int c = 0; c < n_cls; ++c) {
            prefix[(size_t)i * (n_cls + 1) + c] = (int)slots.size() - off[i];
            std::sort(bucket[c].begin(), bucket[c].end());
            slots.insert(slots.end(), bucket[c].begin(), bucket[c].end());
        }
        prefix[(size_t)i * (n_cls + 1) + n_cls] = L;
        h->scr_len[i] = L;
    }
    off[n_res] = (int)slots.size();
    if (getenv("MCSCE_B200_SCREEN_DEBUG")) {
        long long sl = 0, sa = 0; int ns = 0;
        for (int i = 0; i < n_res; ++i) if (h->scr_len[i] > 0) {
            ++ns; sl += h->scr_len[i];
            for (int c = 0; c < n_cls; ++c) sa += h->cls_active[(size_t)i * n_cls + c];
        }
        fprintf(stderr, "[mcsce_b200] neighbour screen: %d residues screened, mean list %.0f of %.0f active atoms\n", ns,
                ns ? (double)sl / ns : 0.0, ns ? (double)sa / ns : 0.0);
    }
    if (slots.empty()) slots.push_back(0);
    return 0;
}

extern "C" int mcsce_set_backbone(mcsce_handle h, const mcsce_backbone_desc *b) {
    if (!h || !b) return fail("null argument");
    if (!h->has_sys) return fail("mcsce_set_system must be called first");
    CK(cudaSetDevice(h->device));
    const bool dbg = getenv("MCSCE_B200_SCREEN_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    auto t_start = now();
    CK(cudaDeviceSynchronize());          // a growth still in flight reads these tables
    if (dbg) fprintf(stderr, "[mcsce_b200] set_backbone: sync %.2f ms\n", ms_since(t_start));
    h->table_version++;
    SysDev &d = h->d;
    d.n_trials_total = b->n_trials_total;
    h->max_trials = 0; h->max_trial_atoms = 0;
    for (int i = 0; i < d.n_res; ++i) {
        StepDev &o = h->steps[i];
        o.n_trials = b->step_n_trials[i]; o.trial_off = b->step_trial_off[i];
        o.n_chi_lib = b->step_n_chi_lib[i]; o.n_chi = b->step_n_chi[i];
        if (o.n_chi > MAX_ROT || o.n_chi_lib > MAX_CHI) return fail("too many chi angles");
        for (int k = 0; k < 4; ++k) { o.q1[k] = b->place_q1[(size_t)i * 4 + k]; o.q2[k] = b->place_q2[(size_t)i * 4 + k]; }
        for (int k = 0; k < 3; ++k) o.ca[k] = b->place_ca[(size_t)i * 3 + k];
        o.ca[3] = 0.f;
        o.n_var = 0; o.n_fix = 0;
        if (o.kind == 2) {
            const TmplDev &tm = h->tmpl_host[o.type];
            for (int a = 0; a < tm.n_atoms; ++a) {
                const int k = tm.sc_rank[a];
                if (k < 0) continue;
                bool moves = false;
                for (int r = 0; r < o.n_chi; ++r) if ((tm.move[r] >> a) & 1u) moves = true;
                if (o.n_chi > 0 && a == tm.axis_to[0]) moves = false;       // pivot of the first rotation: stays put exactly
                if (moves) o.var_k[o.n_var++] = (unsigned char)k; else o.fix_k[o.n_fix++] = (unsigned char)k;
            }
            // keep the reference's atom order inside each list
            std::sort(o.var_k, o.var_k + o.n_var); std::sort(o.fix_k, o.fix_k + o.n_fix);
        }
        if (o.kind == 2) {
            h->max_trials = std::max(h->max_trials, o.n_trials);
            h->max_trial_atoms = std::max(h->max_trial_atoms, o.n_trials * o.n_sc);
        }
    }
    if (ensure_dyn_smem(select_kernel, sizeof(double) * 2 * (size_t)h->max_trials, "mcsce_set_backbone")) return -1;
    std::vector<int> aoff((size_t)d.n_res + 1, 0);       // persistent growth: where a residue's placed trials start
    h->trial_atoms_total = 0;
    for (int i = 0; i < d.n_res; ++i) {
        const StepDev &o = h->steps[i];
        aoff[i] = (int)std::min<long long>(h->trial_atoms_total, 0x7fffffffLL);
        if (o.kind == 2 && o.n_trials > 0) h->trial_atoms_total += (long long)o.n_trials * o.n_sc;
    }
    aoff[d.n_res] = (int)std::min<long long>(h->trial_atoms_total, 0x7fffffffLL);
    h->input_xyz_host.assign(b->input_xyz, b->input_xyz + (size_t)d.n_input * 3);
    std::vector<int> scr_off, scr_slots, scr_prefix;
    if (build_screen_lists(h, b, scr_off, scr_slots, scr_prefix)) return -1;
    if (dbg) fprintf(stderr, "[mcsce_b200] set_backbone: +screen lists (host) %.2f ms\n", ms_since(t_start));
    // All backbone-dependent device tables live in one grow-only allocation and travel in ONE copy from a pinned staging
    // buffer (no cudaMalloc/cudaFree per backbone, no per-table synchronous copies); the kernels that finish the tables
    // (GLY/ALA placement, input-atom energy) follow on the same stream - nothing here waits for the device again.
    float *rigid = nullptr;
    {
        struct Item { const void *src; size_t bytes; const void **dst; };
        const size_t nt6 = (size_t)b->n_trials_total * MAX_CHI;
        const void *steps_dst = nullptr, *rigid_dst = nullptr;
        const Item items[] = {
            {h->steps.data(), sizeof(StepDev) * (size_t)d.n_res, &steps_dst},
            {b->chi_mean, nt6 * sizeof(double), (const void **)&d.chi_mean},
            {b->chi_sigma, nt6 * sizeof(double), (const void **)&d.chi_sigma},
            {b->prob, (size_t)b->n_trials_total * sizeof(double), (const void **)&d.prob},
            {b->input_xyz, (size_t)d.n_input * 3 * sizeof(float), (const void **)&d.input_xyz},
            {scr_off.data(), scr_off.size() * sizeof(int), (const void **)&d.scr_off},
            {scr_slots.data(), scr_slots.size() * sizeof(int), (const void **)&d.scr_slot},
            {scr_prefix.data(), scr_prefix.size() * sizeof(int), (const void **)&d.scr_prefix},
            {aoff.data(), aoff.size() * sizeof(int), (const void **)&h->d_aoff},
            {nullptr, (size_t)d.n_atoms * 3 * sizeof(float), &rigid_dst},                  // rigid_xyz: NaN = not a rigid atom
            {nullptr, sizeof(double2) * (size_t)std::max(1, d.n_input), (const void **)&h->d_bb_rows},    // input-energy scratch
            {nullptr, sizeof(double2) * (size_t)std::max(1, d.n_bb_pairs), (const void **)&h->d_bb_pairs},
        };
        size_t need = 0;
        for (const Item &it : items) need += (std::max<size_t>(it.bytes, 1) + 255) / 256 * 256;
        if (need > h->bb_cap) {
            if (h->bb_slab) cudaFree(h->bb_slab);
            if (h->bb_stage) cudaFreeHost(h->bb_stage);
            h->bb_slab = nullptr; h->bb_stage = nullptr; h->bb_cap = 0;
            CK(cudaMalloc((void **)&h->bb_slab, need + need / 4));
            CK(cudaMallocHost((void **)&h->bb_stage, need + need / 4));
            h->bb_cap = need + need / 4;
        }
        size_t used = 0, copy_bytes = 0;
        for (const Item &it : items) {
            char *p = h->bb_slab + used;
            if (it.src) { memcpy(h->bb_stage + used, it.src, it.bytes); }
            else if (it.dst == &rigid_dst) {
                float *f = reinterpret_cast<float *>(h->bb_stage + used);
                for (size_t k = 0; k < (size_t)d.n_atoms * 3; ++k) f[k] = NAN;
            }
            used += (std::max<size_t>(it.bytes, 1) + 255) / 256 * 256;
            if (it.src || it.dst == &rigid_dst) copy_bytes = used;
            *it.dst = p;
        }
        CK(cudaMemcpyAsync(h->bb_slab, h->bb_stage, copy_bytes, cudaMemcpyHostToDevice, 0));
        d.steps = reinterpret_cast<const StepDev *>(steps_dst);
        d.rigid_xyz = reinterpret_cast<const float *>(rigid_dst);
        rigid = const_cast<float *>(d.rigid_xyz);
    }
    if (dbg) fprintf(stderr, "[mcsce_b200] set_backbone: +tables %.2f ms\n", ms_since(t_start));
    place_rigid_kernel<<<std::max(1, d.n_res), 32>>>(d, rigid);
    h->launches += 1;
    // energy of the input atoms alone (side_chain_builder.py:149): a constant of the backbone, every conformer starts from it
    if (launch_input_energy(h, const_cast<double2 *>(h->d_bb_rows), const_cast<double2 *>(h->d_bb_pairs), h->d_input_energy, 0)) return -1;
    CK(cudaEventRecord(h->bb_ready, 0));      // growths on other streams wait for the tables (mcsce_grow)
    if (dbg) fprintf(stderr, "[mcsce_b200] set_backbone: +rigid side chains, input energy enqueued %.2f ms\n", ms_since(t_start));
    h->has_bb = true;
    return 0;
}

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

static int trials_per_tile_for(int n_per) { return std::max(1, SCREEN_TILE_A / std::max(1, n_per)); }   // screen pass
// tiles the scoring pass can need for up to n_trials surviving trials (the device picks the layout per conformer)
static int max_tiles_for(TileRule rule, int n_trials, int n_per, int n_fix, bool any_count) {
    int best = 1;
    for (int nk = any_count ? 1 : std::max(1, n_trials); nk <= n_trials; ++nk) {
        int n, t;
        tile_layout(rule, nk, n_per, n_fix, n, t);
        best = std::max(best, n);
    }
    return best;
}

// Environment split of the scoring pass: 1 when the conformers (x trial tiles) fill the GPU, otherwise enough shares of the
// environment to reach about one resident round of CTAs.  (A rule that also balanced the last round of big launches
// was measured: no gain on cfg3/cfg5, and it makes the float32 partial sums depend on the number of conformers per GPU.)
static int choose_split(const mcsce_handle h, int step, int n_sets, int n_trials, int forced, bool screened = false) {
    const StepDev &st = h->steps[step];
    int n_active = 0;
    for (int c = 0; c < h->d.n_cls; ++c) n_active += h->cls_active[(size_t)step * h->d.n_cls + c];
    if (screened) n_trials = std::max(1, (n_trials * 5 + 7) / 8);      // the screen leaves about 5 trials in 8 (measured, cfg3/cfg5)
    int tiles, tpt_;
    tile_layout(h->rule, n_trials, st.n_var > 0 && st.n_fix > 0 ? st.n_var : st.n_sc, st.n_var > 0 ? st.n_fix : 0, tiles, tpt_);
    const long long ctas = (long long)tiles * n_sets;
    int max_split = std::max(1, std::min(32, (n_active + 255) / 256));
    if (forced > 0) return std::min(forced, max_split);
    // about two resident rounds of CTAs: a launch ends with a tail of one CTA's duration in which the SMs run dry, and smaller
    // CTAs shorten it (measured, config 3 at 256 / 512 conformers: 87.9 -> 84.2 ms / 155.1 -> 152.0 ms against one round)
    const long long want = 8LL * h->sm_count;
    int split = (int)std::min<long long>(max_split, std::max<long long>(1, (want + ctas / 2) / ctas));
    return split;
}
static int max_split_for(int n_sets, int sm_count) {
    // split factors the workspace is sized for: 1 when the conformers alone fill the GPU, else the maximum (32),
    // because the factor is raised while conformers die (see mcsce_grow)
    long long want = 8LL * sm_count;
    return (want + n_sets / 2) / n_sets > 1 ? 32 : 1;
}

struct Workspace {
    float4 *env, *trial, *trial2;
    double2 *partial, *sp_partial;
    unsigned char *alive, *screen;
    int *n_keep;
    double *e_acc;
    size_t total;
};

static Workspace carve(const mcsce_handle h, int n_conf, int max_trials, int max_trial_atoms, void *base) {
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return (char *)base + o; };
    const int ms = max_split_for(n_conf, h->sm_count);
    w.env = (float4 *)take((size_t)n_conf * h->d.n_atoms * sizeof(float4));
    w.trial = (float4 *)take((size_t)n_conf * std::max(1, max_trial_atoms) * sizeof(float4));
    w.trial2 = (float4 *)take((size_t)n_conf * std::max(1, max_trial_atoms) * sizeof(float4));
    w.partial = (double2 *)take((size_t)n_conf * (std::max(1, max_trials) + 1) * ms * sizeof(double2));
    w.sp_partial = (double2 *)take((size_t)n_conf * std::max(1, max_trials) * sizeof(double2));
    w.alive = (unsigned char *)take((size_t)n_conf);
    w.screen = (unsigned char *)take((size_t)n_conf * std::max(1, max_trials));
    w.n_keep = (int *)take((size_t)n_conf * sizeof(int));
    w.e_acc = (double *)take((size_t)n_conf * sizeof(double));
    w.total = off;
    return w;
}

static int lane_count(const mcsce_handle h, int n_conf);
static size_t gc_workspace_bytes(const mcsce_handle h, int C);
// Launch-bound runs: little pair work per residue (few conformers and/or a small protein), where launch latencies and tails
// of the per-kernel path outweigh the arithmetic.
static bool launch_bound(const mcsce_handle h, int C, long long *grow_steps_out = nullptr, long long *pairs_per_step_out = nullptr) {
    const SysDev &d = h->d;
    long long grow_steps = 0, pairs = 0;
    for (int i = 0; i < d.n_res; ++i) {
        const StepDev &sd = h->steps[i];
        if (sd.kind != 2 || sd.n_trials <= 0) continue;
        long long n_active = 0;
        for (int c = 0; c < d.n_cls; ++c) n_active += h->cls_active[(size_t)i * d.n_cls + c];
        pairs += (long long)sd.n_trials * sd.n_sc * n_active;
        ++grow_steps;
    }
    if (grow_steps_out) *grow_steps_out = grow_steps;
    if (pairs_per_step_out) *pairs_per_step_out = grow_steps > 0 ? pairs / grow_steps : 0;      // per conformer
    return grow_steps > 0 && (pairs * C) / grow_steps < 300000000LL;
}
static int persistent_knob() {               // MCSCE_B200_PERSISTENT: 0 = never picked automatically, 1 = launch-bound runs (default), 2 = always
    if (const char *e = getenv("MCSCE_B200_PERSISTENT")) return atoi(e);
    return 1;
}
#define GC_MAX_WORKSPACE (2048LL << 20)      // the up-front placement is for launch-bound runs: beyond 2 GB the per-kernel path serves

extern "C" size_t mcsce_workspace_bytes(mcsce_handle h, int n_conf) {
    if (!h || !h->has_bb || n_conf <= 0) return 0;
    size_t whole = carve(h, n_conf, h->max_trials, h->max_trial_atoms, nullptr).total;
    const int G = lane_count(h, n_conf);
    if (G > 1) {                             // conformer groups, each with a complete workspace
        size_t sum = 0;
        for (int g = 0; g < G; ++g) {
            const int c0 = (int)((long long)n_conf * g / G), c1 = (int)((long long)n_conf * (g + 1) / G);
            sum += carve(h, c1 - c0, h->max_trials, h->max_trial_atoms, nullptr).total;
        }
        whole = std::max(whole, sum);
    }
    // the persistent cluster growth (launch-bound runs) places every trial up front
    const size_t pers = gc_workspace_bytes(h, n_conf);
    if (pers <= (size_t)GC_MAX_WORKSPACE && (launch_bound(h, n_conf) || persistent_knob() >= 2)) whole = std::max(whole, pers);
    return whole;
}

static int launch_input_energy(mcsce_handle h, double2 *rows, double2 *pairs, double *out, cudaStream_t st) {
    const SysDev &d = h->d;
    input_energy_rows_kernel<<<(d.n_input + 127) / 128, 128, 0, st>>>(d, rows);
    if (d.n_bb_pairs) input_energy_pairs_kernel<<<(d.n_bb_pairs + 127) / 128, 128, 0, st>>>(d, pairs);
    input_energy_reduce_kernel<<<1, 256, 0, st>>>(d.n_input, rows, d.n_bb_pairs, pairs, out);
    h->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int mcsce_input_energy(mcsce_handle h, double *d_energy, void *stream) {
    if (!h || !h->has_bb) return fail("set_system/set_backbone first");
    CK(cudaSetDevice(h->device));
    // the value is a constant of the backbone, computed by set_backbone: hand out a copy
    CK(cudaStreamWaitEvent((cudaStream_t)stream, h->bb_ready, 0));
    CK(cudaMemcpyAsync(d_energy, h->d_input_energy, sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int mcsce_init_env(mcsce_handle h, int n_env_sets, float *d_env, void *stream) {
    if (!h || !h->has_bb) return fail("set_system/set_backbone first");
    CK(cudaSetDevice(h->device));
    if (!h->capturing) CK(cudaStreamWaitEvent((cudaStream_t)stream, h->bb_ready, 0));
    const long long n = (long long)n_env_sets * h->d.n_atoms;
    init_env_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d, n_env_sets, (float4 *)d_env);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

static int launch_place(mcsce_handle h, int step, int n_sets, int n_rows, const double *chis_in, long long set_stride,
                        long long row_off, double *chis_out, float4 *trial, long long trial_stride,
                        const unsigned char *alive, cudaStream_t st, const double *ori_in) {
    PlaceArgs a{};
    a.ori_in = ori_in;
    a.s = h->d; a.step = step; a.n_trials = n_rows; a.chis_in = chis_in; a.chis_set_stride = set_stride;
    a.chis_row_off = row_off; a.chis_out = chis_out; a.trial = trial; a.trial_set_stride = trial_stride;
    a.alive = alive; a.cp = h->cur_call;
    dim3 grid((n_rows + K1_TRIALS - 1) / K1_TRIALS, n_sets);
    { ProfScope ps(h, 0, st); place_trials_kernel<<<grid, 128, 0, st>>>(a); }
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

static int launch_generic(mcsce_handle h, int step, int n_sets, int n_trials, int n_split, const float4 *env,
                          const float4 *trial, long long trial_stride, double2 *partial, const unsigned char *alive,
                          cudaStream_t st, int dedup = 0, unsigned char *screen = nullptr, const double2 *sp_partial = nullptr,
                          int *n_keep = nullptr, int call_sets = 0) {
    const StepDev &sd = h->steps[step];
    if (dedup && (sd.n_fix <= 0 || sd.n_var <= 0)) dedup = 0;
    ScoreArgs a{};
    a.s = h->d; a.step = step; a.n_trials = n_trials; a.trials_per_tile = trials_per_tile_for(dedup ? sd.n_var : sd.n_sc);
    a.n_split = n_split; a.env = env; a.trial = trial; a.trial_set_stride = trial_stride; a.partial = partial; a.alive = alive;
    a.dedup = dedup; a.screen = screen; a.sp_partial = sp_partial; a.n_keep = n_keep;
    const int tiles = (n_trials + a.trials_per_tile - 1) / a.trials_per_tile;        // screen pass: fixed 256-atom tiles
    const int n_per = dedup ? sd.n_var : sd.n_sc;
    a.rule = h->rule;
    dim3 grid(max_tiles_for(h->rule, n_trials, n_per, dedup ? sd.n_fix : 0, screen != nullptr), n_sets, n_split);
    {
        ProfScope ps(h, 1, st);
        // scoring pass: one CTA per conformer and environment share (score_conf.cuh) for big ensembles - at least 1024
        // conformers in the call (all conformer groups together) and an environment split of at most 2 -, whose share fits two
        // such CTAs per SM; else one CTA per 256-atom trial tile.  Measured on config 3 (B200, ms per ensemble, per-conformer /
        // per-tile CTAs): 2048 conformers 520.1 / 536.7, 1024: 260.5 / 269.5, 512: 138.4 / 139.3.  Its CTAs are 3-20x coarser, so
        // it loses where CTAs are few or die unevenly (256 conformers 82.0 / 74.9; config 4 x 512, 40 % of the conformers fail:
        // 33.7 / 31.4; config 5 x 32, split 32: 393 / 333) and in the launch-bound regime (76 residues x 128 as a graph: 5.20 / 4.55).
        // MCSCE_B200_K2CONF: 0 = never, 1 = that rule (default), 2 = whenever the share fits.  With an environment split of 1 the
        // neighbour screen runs inside that kernel (MCSCE_B200_K2CONF_FUSE=0: as a launch of its own).
        if (h->k2conf < 0) { const char *e = getenv("MCSCE_B200_K2CONF"); h->k2conf = e ? atoi(e) : 1; }
        bool use_conf = false, fuse = false;
        int cap_pairs = 0, list_cap_pairs = 0;
        size_t smem = 0;
        if (h->k2conf >= 2 || (h->k2conf == 1 && call_sets >= 1024 && n_split <= 2)) {
            int n_active = 0;
            for (int c = 0; c < h->d.n_cls; ++c) n_active += h->cls_active[(size_t)step * h->d.n_cls + c];
            const int per = (n_active + n_split - 1) / n_split;
            cap_pairs = (per + h->d.n_cls + 2) / 2 + 1;
            static int fuse_knob = -1;
            if (fuse_knob < 0) { const char *e = getenv("MCSCE_B200_K2CONF_FUSE"); fuse_knob = e ? atoi(e) : 1; }
            fuse = screen != nullptr && n_split == 1 && fuse_knob > 0;
            if (fuse) list_cap_pairs = (h->scr_len[step] + h->d.n_cls + 2) / 2 + 1;
            smem = (size_t)(cap_pairs + list_cap_pairs) * 32 + (size_t)SC_SE * 9 + (size_t)SCREEN_WORDS * 32;
            use_conf = smem <= (size_t)110 * 1024;
        }
        if (screen && !(use_conf && fuse)) {      // neighbour screen first: flags the trials the scoring pass can skip
            score_generic_kernel<true><<<dim3(tiles, n_sets, 1), SCORE_THREADS, 0, st>>>(a);
            h->launches += 1;
        }
        bool done = false;
        if (use_conf) {
            if (h->k2conf_smem == 0) {
                CK(cudaFuncSetAttribute(score_conf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
                h->k2conf_smem = (size_t)110 * 1024;
            }
            static int fixed_shares = -1;
            // four fixed shares of the staged pairs per 64-atom chunk: the per-atom sums do not depend on how many trials
            // the screen left (bit-identical with and without it); 0 = as many as fill the warps best (0.5 % faster)
            if (fixed_shares < 0) { const char *e = getenv("MCSCE_B200_K2CONF_E"); fixed_shares = e ? atoi(e) : 4; }
            score_conf_kernel<<<dim3(n_sets, n_split), SC_THREADS, smem, st>>>(a, cap_pairs, fixed_shares, fuse ? list_cap_pairs : 0);
            done = true;
        }
        if (!done) score_generic_kernel<false><<<grid, SCORE_THREADS, 0, st>>>(a);
    }
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

static int launch_special(mcsce_handle h, int step, int n_sets, int n_trials, const float4 *env, const float4 *trial,
                          long long trial_stride, double2 *sp_partial, const unsigned char *alive, cudaStream_t st) {
    SpecialArgs b{};
    b.s = h->d; b.step = step; b.n_trials = n_trials; b.env = env; b.trial = trial; b.trial_set_stride = trial_stride;
    b.sp_partial = sp_partial; b.alive = alive;
    dim3 grid2((n_trials + 8 * SP_TPW - 1) / (8 * SP_TPW), n_sets);
    { ProfScope ps(h, 2, st); score_special_kernel<<<grid2, 256, 0, st>>>(b); }
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

static int launch_score(mcsce_handle h, int step, int n_sets, int n_trials, int n_split, const float4 *env,
                        const float4 *trial, long long trial_stride, double2 *partial, double2 *sp_partial,
                        const unsigned char *alive, cudaStream_t st) {
    if (launch_generic(h, step, n_sets, n_trials, n_split, env, trial, trial_stride, partial, alive, st)) return -1;
    return launch_special(h, step, n_sets, n_trials, env, trial, trial_stride, sp_partial, alive, st);
}

extern "C" int mcsce_place_trials_replay(mcsce_handle h, int residue, int n_rows, const double *d_chis, const double *d_ori,
                                         float *d_xyz, void *stream) {
    if (!h || !h->has_bb) return fail("set_system/set_backbone first");
    if (residue < 0 || residue >= h->d.n_res) return fail("residue out of range");
    if (h->steps[residue].kind == 0) return fail("residue is retained: nothing to place");
    if (!d_chis) return fail("mcsce_place_trials: chi rows are required");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamWaitEvent((cudaStream_t)stream, h->bb_ready, 0));
    return launch_place(h, residue, 1, n_rows, d_chis, 0, 0, nullptr, (float4 *)d_xyz, 0, nullptr, (cudaStream_t)stream, d_ori);
}

extern "C" int mcsce_place_trials(mcsce_handle h, int residue, int n_rows, const double *d_chis, float *d_xyz, void *stream) {
    return mcsce_place_trials_replay(h, residue, n_rows, d_chis, nullptr, d_xyz, stream);
}

__global__ void combine_kernel(int n, int n_split, const double2 *partial, const double2 *sp_partial, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e = sp_partial[i].x; bool cl = sp_partial[i].y != 0.0;
    for (int z = 0; z < n_split; ++z) { e += partial[(long long)i * n_split + z].x; cl |= partial[(long long)i * n_split + z].y != 0.0; }
    out[i] = cl ? CUDART_INF : e;
}

extern "C" int mcsce_score_trials(mcsce_handle h, int residue, int n_env_sets, int trials_per_env, const float *d_trial,
                                  const float *d_env, double *d_energy, void *d_workspace, size_t workspace_bytes,
                                  int env_split, void *stream) {
    if (!h || !h->has_bb) return fail("set_system/set_backbone first");
    if (residue < 0 || residue >= h->d.n_res || h->steps[residue].kind != 2) return fail("residue is not a grown residue");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamWaitEvent((cudaStream_t)stream, h->bb_ready, 0));
    const int split = choose_split(h, residue, n_env_sets, trials_per_env, env_split);
    const size_t n = (size_t)n_env_sets * trials_per_env;
    const size_t need = align_up(n * split * sizeof(double2)) + align_up(n * sizeof(double2));
    if (workspace_bytes < need) return fail("workspace too small for mcsce_score_trials");
    double2 *partial = (double2 *)d_workspace;
    double2 *sp = (double2 *)((char *)d_workspace + align_up(n * split * sizeof(double2)));
    cudaStream_t st = (cudaStream_t)stream;
    const long long stride = (long long)trials_per_env * h->steps[residue].n_sc;
    if (launch_score(h, residue, n_env_sets, trials_per_env, split, (const float4 *)d_env, (const float4 *)d_trial, stride,
                     partial, sp, nullptr, st)) return -1;
    combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((int)n, split, partial, sp, d_energy);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" size_t mcsce_score_workspace_bytes(mcsce_handle h, int n_env_sets, int trials_per_env) {
    const size_t n = (size_t)n_env_sets * trials_per_env;
    return align_up(n * 32 * sizeof(double2)) + align_up(n * sizeof(double2));
}

__global__ void pack_env_kernel(SysDev s, int n_sets, int n_old, const float *xyz, float4 *env) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_sets * s.n_atoms) return;
    const int atom = (int)(i % s.n_atoms);
    const int set = (int)(i / s.n_atoms);
    float4 v = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.f);
    if (atom < n_old) {
        const float *p = xyz + ((long long)set * n_old + atom) * 3;
        v = make_float4(p[0], p[1], p[2], (float)s.charge[atom]);
    }
    env[(long long)set * s.n_atoms + s.slot_of_atom[atom]] = v;
}

extern "C" int mcsce_pack_env(mcsce_handle h, int n_env_sets, int n_old, const float *d_old_xyz, float *d_env, void *stream) {
    if (!h || !h->has_sys) return fail("set_system first");
    if (n_old < 0 || n_old > h->d.n_atoms) return fail("n_old out of range");
    CK(cudaSetDevice(h->device));
    const long long n = (long long)n_env_sets * h->d.n_atoms;
    pack_env_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d, n_env_sets, n_old, d_old_xyz, (float4 *)d_env);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int mcsce_select(mcsce_handle h, int n_sets, int n_trials, const double *d_energy, const double *d_prob,
                            double beta, const double *d_uniform, int32_t *d_selected, double *d_e_selected, void *stream) {
    if (!h) return fail("null handle");
    if (n_sets <= 0 || n_trials <= 0) return fail("mcsce_select: n_sets and n_trials must be positive");
    CK(cudaSetDevice(h->device));
    if (ensure_dyn_smem(select_plain_kernel, sizeof(double) * (size_t)n_trials, "mcsce_select")) return -1;
    select_plain_kernel<<<n_sets, 128, sizeof(double) * n_trials, (cudaStream_t)stream>>>(n_trials, d_energy, d_prob, beta,
                                                                                         d_uniform, d_selected, d_e_selected);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// Per grown residue, the atoms already placed that can come within `range` of any of its side-chain atoms, whatever the chi
// angles of either: input atoms by their position, side-chain atoms of earlier residues by their CA and reach bound.  Complete
// (unlike the clash screen's lists, which only have to be likely): the hydrophobic PMF visits no other atom for its surface terms.
// Upper bound of |atom - CA| over all chi angles, per template atom, for the first n_rot rotations of the template (the ones
// a residue applies).  Rotation k turns its moving set about the axis (axis_from[k] -> axis_to[k]) through the pivot
// axis_to[k]; atoms on the axis stay put.  Every other rotation moves an atom and the pivot of the last rotation that moves it
// together, so |atom - pivot| is a template constant and  bound(atom) <= bound(pivot) + |atom - pivot|.  An axis through CA
// (chi1) keeps the distance to CA itself.  2 % + 0.05 A of slack for the float32 arithmetic of K1.
static void reach_bounds(const TmplDev &o, int n_rot, float *out) {
    auto dist = [&](int a, int b) {
        const float x = o.xyz[a][0] - o.xyz[b][0], y = o.xyz[a][1] - o.xyz[b][1], z = o.xyz[a][2] - o.xyz[b][2];
        return sqrtf(x * x + y * y + z * z);
    };
    auto norm = [&](int a) { return sqrtf(o.xyz[a][0] * o.xyz[a][0] + o.xyz[a][1] * o.xyz[a][1] + o.xyz[a][2] * o.xyz[a][2]); };
    float far = 0.f;
    for (int a = 0; a < o.n_atoms; ++a) far = std::max(far, norm(a));
    std::vector<float> bound(MAX_TMPL, -1.f);
    std::function<float(int, int)> rec = [&](int a, int depth) -> float {
        if (bound[a] >= 0.f) return bound[a];
        int m = -1;
        for (int k = 0; k < n_rot; ++k) if (((o.move[k] >> a) & 1u) && a != o.axis_to[k] && a != o.axis_from[k]) m = k;
        float b;
        if (m < 0) b = norm(a);
        else if (norm(o.axis_from[m]) < 1e-4f || norm(o.axis_to[m]) < 1e-4f) b = norm(a);       // the axis passes through CA
        else if (depth > MAX_ROT + 2) return 3.f * far + 3.f;       // rotation programs that are not nested chains: a loose bound
        else b = rec(o.axis_to[m], depth + 1) + dist(a, o.axis_to[m]);
        return bound[a] = b;
    };
    for (int a = 0; a < MAX_TMPL; ++a) out[a] = a < o.n_atoms ? 1.02f * rec(a, 0) + 0.05f : 0.f;
}

// Per grown residue, the atoms already placed that can come within `range` of any of its side-chain atoms, whatever the chi
// angles of either: input atoms by their position, side-chain atoms of earlier residues by their CA and reach bound.  Complete
// (unlike the clash screen's lists, which only have to be likely): the hydrophobic PMF visits no other atom for its surface terms.
static int ensure_hpmf_near_lists(mcsce_handle h, mcsce_hpmf hp) {
    if (mcsce_hpmf_near_version(hp) == h->table_version) return 0;
    const SysDev &d = h->d;
    if (mcsce_hpmf_n_atoms(hp) != d.n_atoms) return fail("the hpmf handle was created for another atom list");
    const double range = mcsce_hpmf_sasa_range(hp) + 0.01;
    std::vector<float> cx((size_t)d.n_atoms * 3, 0.f), rad(d.n_atoms, 0.f), reach(d.n_res, 0.f);
    for (int a = 0; a < d.n_input; ++a) for (int k = 0; k < 3; ++k) cx[(size_t)a * 3 + k] = h->input_xyz_host[(size_t)a * 3 + k];
    float bound[MAX_TMPL];
    for (int j = 0; j < d.n_res; ++j) {
        const StepDev &sj = h->steps[j];
        if (sj.kind == 0) continue;
        const TmplDev &tm = h->tmpl_host[sj.type];
        reach_bounds(tm, sj.kind == 2 ? std::min(sj.n_chi, tm.n_chi) : 0, bound);
        for (int p = 0; p < tm.n_atoms; ++p) {
            const int k = tm.sc_rank[p];
            if (k < 0 || k >= sj.n_sc) continue;
            const int a = sj.sc_start + k;
            for (int q = 0; q < 3; ++q) cx[(size_t)a * 3 + q] = sj.ca[q];
            rad[a] = bound[p];
            reach[j] = std::max(reach[j], rad[a]);
        }
    }
    const double range2 = mcsce_hpmf_pair_range(hp) + 0.01;       // carbon pairs: the Gaussian cutoff
    h->hp_reach = reach;
    std::vector<int> off(d.n_res + 1, 0), atoms, off2(d.n_res + 1, 0), atoms2;
    for (int i = 0; i < d.n_res; ++i) {
        off[i] = (int)atoms.size(); off2[i] = (int)atoms2.size();
        const StepDev &si = h->steps[i];
        if (si.kind != 2) continue;
        for (int a = 0; a < si.sc_start; ++a) {
            const float dx = cx[(size_t)a * 3] - si.ca[0], dy = cx[(size_t)a * 3 + 1] - si.ca[1], dz = cx[(size_t)a * 3 + 2] - si.ca[2];
            const double d2 = (double)dx * dx + (double)dy * dy + (double)dz * dz;
            const double lim = (double)reach[i] + rad[a] + range, lim2 = (double)reach[i] + rad[a] + range2;
            if (d2 < lim * lim) atoms.push_back(a);
            if (d2 < lim2 * lim2) atoms2.push_back(a);
        }
    }
    off[d.n_res] = (int)atoms.size(); off2[d.n_res] = (int)atoms2.size();
    if (atoms.empty()) atoms.push_back(0);
    if (atoms2.empty()) atoms2.push_back(0);
    return mcsce_hpmf_set_near_lists(hp, h->table_version, d.n_res, off.data(), atoms.data(), off2.data(), atoms2.data());
}

static int ensure_hpmf_buffers(mcsce_handle h, mcsce_handle_s::Lane &ln, int n_conf, cudaStream_t st) {
    const size_t na = (size_t)h->d.n_atoms, nt = (size_t)std::max(1, h->max_trials);
    if ((size_t)n_conf <= ln.hp_cap_conf && na <= ln.hp_cap_atoms && nt <= ln.hp_cap_trials) return 0;
    CK(cudaDeviceSynchronize());
    if (ln.hp_gcoords) cudaFree(ln.hp_gcoords);
    if (ln.hp_v) cudaFree(ln.hp_v);
    if (ln.hp_mask) cudaFree(ln.hp_mask);
    if (ln.hp_ovf) cudaFree(ln.hp_ovf);
    ln.hp_gcoords = nullptr; ln.hp_v = nullptr; ln.hp_mask = nullptr; ln.hp_ovf = nullptr; ln.hp_cap_conf = 0;
    const size_t C = std::max<size_t>(n_conf, ln.hp_cap_conf), A = std::max(na, ln.hp_cap_atoms), T = std::max(nt, ln.hp_cap_trials);
    CK(cudaMalloc((void **)&ln.hp_gcoords, C * A * 3 * sizeof(float)));
    CK(cudaMalloc((void **)&ln.hp_v, C * T * sizeof(double)));
    CK(cudaMalloc((void **)&ln.hp_mask, C * T));
    CK(cudaMalloc((void **)&ln.hp_ovf, C * sizeof(int)));
    if (!ln.hp_v_init) CK(cudaMalloc((void **)&ln.hp_v_init, sizeof(double)));
    ln.hp_cap_conf = C; ln.hp_cap_atoms = A; ln.hp_cap_trials = T;
    (void)st;
    return 0;
}

__global__ void gather_kernel(SysDev s, int n_sets, const float4 *env, float *coords);

// Enqueue of one growth (environment init, then K1 / K2s / screen / K2 / K3 for every grown residue) on `st` + `aux`, as a
// stepper: begin(), residue(i) for every residue in order, so that several conformer groups (lanes) can be enqueued
// interleaved residue by residue - the occasional read-back of the number of dead conformers (`resync`, never while
// capturing) then stalls the host at the same point of every group's progress instead of serialising the groups.
struct GrowthEnqueue {
    mcsce_handle h; mcsce_grow_opts o; Workspace w; mcsce_handle_s::Lane *ln; cudaStream_t st, aux; bool resync;
    int groups;                 // conformer groups in flight: the environment split is chosen for all of them together
    mcsce_hpmf hp = nullptr;    // coordinate-based term (hydrophobic PMF) of this growth, or null
    int prev = -1, prev2 = -1, count = 0, ms_alloc = 1, n_alive = 0;
    long long trial_stride = 1, chi_set_stride = 0;

    int begin() {
        const SysDev &d = h->d;
        h->cur_call = ln->d_call;
        // per-call scalars -> device parameter block (by-value kernel argument: no host buffer to keep alive)
        // (not while capturing: a replayed graph must read the values of the call that replays it, written by mcsce_grow)
        if (!h->capturing) {
            CallParams cp{o.seed, o.conf_id0, o.beta, o.noise_deg};
            set_call_params_kernel<<<1, 1, 0, st>>>(ln->d_call, cp);
            h->launches += 1;
        }
        const int C = o.n_conf;
        trial_stride = std::max(1, h->max_trial_atoms);
        chi_set_stride = (long long)d.n_trials_total * MAX_CHI;
        if (mcsce_init_env(h, C, (float *)w.env, (void *)st)) return -1;
        hp = (mcsce_hpmf)o.hpmf;
        if (hp) {
            // coordinate-based term: growth-order copy of the conformers for its kernels, its level-0 value (input atoms alone)
            if (ensure_hpmf_near_lists(h, hp) || ensure_hpmf_buffers(h, *ln, C, st)) return -1;
            const long long n = (long long)C * d.n_atoms;
            gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, C, w.env, ln->hp_gcoords);
            h->launches += 1;
            ProfScope ps(h, 4, st);
            if (mcsce_hpmf_growth_begin(hp, (int)(ln - h->lane), ln->hp_gcoords, d.n_input, ln->hp_v_init, st)) return fail(mcsce_last_error());
            h->launches += 4;
        }
        init_state_kernel<<<(C + 255) / 256, 256, 0, st>>>(C, w.alive, w.e_acc, h->d_input_energy, hp ? ln->hp_v_init : nullptr,
                                                           hp ? ln->hp_ovf : nullptr);   // starts from the input atoms' energy (:149)
        h->launches += 1;
        // Two streams: `st` runs the generic-pair kernel and the selection of every residue in order; `aux` runs the
        // placement of the next residue (independent of the pending selection) and the special-pair kernel.
        CK(cudaEventRecord(ln->ev0, st));
        CK(cudaStreamWaitEvent(aux, ln->ev0, 0));
        ms_alloc = max_split_for(C, h->sm_count);
        n_alive = C;
        CK(cudaMemsetAsync(ln->d_n_dead, 0, sizeof(int), st));
        CK(cudaMemsetAsync(w.n_keep, 0, sizeof(int) * (size_t)C, st));
        return 0;
    }

    int residue(int i) {
        const SysDev &d = h->d;
        const StepDev &sd = h->steps[i];
        if (sd.kind != 2 || sd.n_trials <= 0) return 0;
        h->cur_call = ln->d_call;
        const int C = o.n_conf, T = sd.n_trials;
        float4 *trial = (count & 1) ? w.trial2 : w.trial;
        // aux: K1 into the buffer last read by the residue before the previous one
        if (prev2 >= 0) CK(cudaStreamWaitEvent(aux, ln->evS[prev2], 0));
        if (launch_place(h, i, C, T, o.d_chis, chi_set_stride, sd.trial_off, o.d_chis_out, trial, trial_stride,
                         w.alive, aux, o.d_chis ? o.d_ori : nullptr)) return -1;
        CK(cudaEventRecord(ln->evP[i], aux));
        // st: generic pairs
        CK(cudaStreamWaitEvent(st, ln->evP[i], 0));
        // few conformers: the environment is split over blockIdx.z to fill the GPU; conformers that died leave empty
        // CTAs, so the number still alive is read back every 32 residues and the split re-derived from it
        if (resync && ms_alloc > 1 && o.env_split <= 0 && count > 0 && (count & 31) == 0) {
            CK(cudaMemcpyAsync(h->h_n_dead + (ln - h->lane), ln->d_n_dead, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            n_alive = std::max(1, C - h->h_n_dead[ln - h->lane]);
        }
        // neighbour screen: not in the launch-bound (graph) regime, where one more kernel per residue costs more than it saves
        unsigned char *screen = (!o.no_screen && resync && h->scr_len[i] > 0) ? w.screen : nullptr;
        const int dedup = (sd.n_fix > 0 && sd.n_var > 0 && !o.no_dedup) ? 1 : 0;
        const int split = std::min(ms_alloc, choose_split(h, i, n_alive * groups, T, o.env_split, screen != nullptr));
        if (!screen && launch_generic(h, i, C, T, split, w.env, trial, trial_stride, w.partial, w.alive, st, dedup, nullptr, nullptr,
                                      nullptr, resync ? C * groups : 0)) return -1;
        // aux: special pairs (need the environment as updated by the previous selection)
        if (prev >= 0) CK(cudaStreamWaitEvent(aux, ln->evS[prev], 0));
        if (launch_special(h, i, C, T, w.env, trial, trial_stride, w.sp_partial, w.alive, aux)) return -1;
        CK(cudaEventRecord(ln->evQ[i], aux));
        // screened residue: the special pairs go first, their clash flags join the screen's, then the scoring pass
        if (screen) {
            CK(cudaStreamWaitEvent(st, ln->evQ[i], 0));
            if (launch_generic(h, i, C, T, split, w.env, trial, trial_stride, w.partial, w.alive, st, dedup, screen, w.sp_partial,
                               w.n_keep, resync ? C * groups : 0)) return -1;
        }
        // st: selection
        CK(cudaStreamWaitEvent(st, ln->evQ[i], 0));
        SelectArgs a{};
        a.s = d; a.step = i; a.n_trials = T; a.n_split = split; a.partial = w.partial; a.sp_partial = w.sp_partial;
        a.dedup = dedup;
        a.prob = d.prob + sd.trial_off; a.cp = ln->d_call; a.uniform = o.d_uniform; a.u_stride = d.n_res;
        a.env = w.env; a.trial = trial; a.trial_set_stride = trial_stride;
        a.alive = w.alive; a.e_acc = w.e_acc; a.selected = o.d_selected; a.trial_energy = o.d_trial_energy;
        a.pair_counter = h->d_pairs; a.n_dead = ln->d_n_dead;
        long long n_active = 0;
        for (int c = 0; c < d.n_cls; ++c) n_active += h->cls_active[(size_t)i * d.n_cls + c];
        a.pairs_per_set = (long long)T * ((long long)sd.n_sc * (sd.n_sc - 1) / 2 + (long long)sd.n_sc * n_active);
        // executed pairs: chi-independent atoms once, special pairs for every trial, the screen's clash-only pairs at half
        // weight; plus, per trial the screen left, its chi-dependent atoms against the whole environment
        const int n_per = dedup ? sd.n_var : sd.n_sc;
        a.screen = screen; a.n_keep = w.n_keep;
        a.pairs_computed_per_kept = (long long)n_per * n_active;
        a.pairs_computed_per_set = (dedup ? (long long)sd.n_fix * n_active : 0) + (long long)T * sd.n_sp
                                   + (screen ? (long long)T * n_per * h->scr_len[i] / 2 : 0);
        if (hp) {
            // hydrophobic PMF of (conformer so far + trial) for every trial that passed the clash filter (libenergy.py:806-808):
            // the environment's areas / weights / fields are carried from residue to residue, each trial adds its own terms
            a.extra = ln->hp_v; a.gcoords = ln->hp_gcoords; a.term_overflow = ln->hp_ovf;
            ProfScope ps(h, 4, st);
            finite_mask_kernel<<<dim3((T + 127) / 128, C), 128, 0, st>>>(a, ln->hp_mask);
            if (mcsce_hpmf_growth_step(hp, (int)(ln - h->lane), i, C, ln->hp_gcoords, sd.sc_start, sd.n_sc, T, trial, trial_stride,
                                       ln->hp_mask, ln->hp_v, ln->hp_ovf, st, sd.ca, h->hp_reach[i])) return fail(mcsce_last_error());
            h->launches += 5;
        }
        { ProfScope ps(h, 3, st); select_kernel<<<C, 128, sizeof(double) * 2 * T, st>>>(a); }
        h->launches += 1;
        CK(cudaEventRecord(ln->evS[i], st));
        prev2 = prev; prev = i; ++count;
        return 0;
    }
};

static int enqueue_growth(mcsce_handle h, const mcsce_grow_opts *o, const Workspace &w, mcsce_handle_s::Lane &ln,
                          cudaStream_t st, cudaStream_t aux, bool resync) {
    GrowthEnqueue g{h, *o, w, &ln, st, aux, resync, 1};
    if (g.begin()) return -1;
    for (int i = 0; i < h->d.n_res; ++i) if (g.residue(i)) return -1;
    CK(cudaGetLastError());
    return 0;
}

// conformer groups in flight for one call: up to four, each of at least 48 conformers.  The groups share the environment
// split of the whole call (GrowthEnqueue::groups), so a call's result does not depend on the grouping.
static int lane_count(const mcsce_handle h, int n_conf) {
    int g = 1;
    if (const char *e = getenv("MCSCE_B200_LANES")) { g = atoi(e); g = std::max(1, std::min(g, MAX_LANES)); }   // tuning knob
    else g = LANE_LIMIT;
    int per_lane = 48;
    if (const char *e = getenv("MCSCE_B200_LANE_MIN")) per_lane = std::max(1, atoi(e));          // tuning knob
    while (g > 1 && n_conf / g < per_lane) --g;
    return g;
}

static int prepare_lane(mcsce_handle h, mcsce_handle_s::Lane &ln, bool own_main) {
    if (own_main && !ln.main) CK(cudaStreamCreateWithFlags(&ln.main, cudaStreamNonBlocking));
    if (!ln.aux) CK(cudaStreamCreateWithFlags(&ln.aux, cudaStreamNonBlocking));
    if (!ln.ev0) CK(cudaEventCreateWithFlags(&ln.ev0, cudaEventDisableTiming));
    if (!ln.done) CK(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    while ((int)ln.evP.size() < h->d.n_res) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ln.evP.push_back(e);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ln.evQ.push_back(e);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ln.evS.push_back(e);
    }
    return 0;
}

static int finish_growth(mcsce_handle h, int C, const Workspace &w, float *d_coords, double *d_energy, uint8_t *d_ok,
                         cudaStream_t st, const int *term_overflow = nullptr) {
    const long long n = (long long)C * h->d.n_atoms;
    gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d, C, w.env, d_coords);
    finish_kernel<<<(C + 255) / 256, 256, 0, st>>>(C, w.alive, w.e_acc, d_energy, d_ok, term_overflow);
    h->launches += 2;
    CK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------
// Persistent cluster growth (grow_cluster.cuh): launch plan and launch.  Returns 1 when the call was served, 0 when this
// path cannot take it (the caller falls back to the graph / direct path), -1 on error.
// ------------------------------------------------------------------------------------------------
static size_t gc_smem_bytes(const mcsce_handle h) {
    // (mirrors the carve at the top of grow_cluster_kernel)
    const size_t T = (size_t)std::max(1, h->max_trials), Tpad = (T + 1) & ~(size_t)1, n2 = (size_t)h->d.n_cls * h->d.n_cls;
    const size_t b = (size_t)h->env_pairs * 32 + (size_t)GC_SE * 8 + Tpad * 8 * 2 + 2 * Tpad * 8 + 2 * Tpad * 16 + 2 * (Tpad + 2) * 16 +
                     2 * (size_t)GC_TRIAL_PF * 16 + n2 * 16 + ((n2 + 3) & ~(size_t)3) * 4 + GC_SE + (size_t)h->d.n_res;
    return (b + 15) / 16 * 16;
}
// device workspace of the persistent growth: every trial of every residue placed up front, and their special-pair sums
static size_t gc_workspace_bytes(const mcsce_handle h, int C) {
    return align_up((size_t)C * (size_t)std::max<long long>(1, h->trial_atoms_total) * sizeof(float4)) +
           align_up((size_t)C * (size_t)std::max(1, h->d.n_trials_total) * sizeof(double2));
}

static int gc_make_plan(mcsce_handle h, int C) {
    int max_s = 16, thr_override = 0;
    if (const char *e = getenv("MCSCE_B200_CLUSTER_MAX")) max_s = std::max(1, std::min(16, atoi(e)));     // tuning knobs
    if (const char *e = getenv("MCSCE_B200_CLUSTER_THREADS")) thr_override = std::max(64, std::min(GC_MAX_THREADS, atoi(e) / 32 * 32));
    const long long key1 = (long long)C | ((long long)max_s << 32) | ((long long)thr_override << 40);
    if (h->gc_plan_key[0] == h->table_version && h->gc_plan_key[1] == key1) return h->gc_S > 0 ? 1 : 0;
    h->gc_plan_key[0] = h->table_version; h->gc_plan_key[1] = key1; h->gc_S = 0;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, grow_cluster_kernel));
    if (!h->gc_np_set) { CK(cudaFuncSetAttribute(grow_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); h->gc_np_set = true; }
    for (int S = max_s; S >= 1; S >>= 1) {
        if (S > 1 && (long long)C * S > h->sm_count) continue;        // CTAs are added to a conformer only while SMs are spare
        int threads = (long long)C * S <= h->sm_count ? 512 : 256;    // one CTA per SM: 16 warps; otherwise two CTAs of 8
        if (thr_override) threads = thr_override;
        const size_t smem = gc_smem_bytes(h);
        if (smem + fa.sharedSizeBytes > (size_t)227 * 1024) return 0;  // the environment image does not fit one SM
        CK(cudaFuncSetAttribute(grow_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = 0;
        if (S > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(C * S)); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n_active = 0;
            if (cudaOccupancyMaxActiveClusters(&n_active, grow_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
            if (n_active < C) continue;                                // every conformer's cluster must be resident at once
            grid = C * S;
        } else {
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grow_cluster_kernel, threads, smem));
            if (per_sm < 1) return 0;
            grid = (int)std::min<long long>(C, (long long)per_sm * h->sm_count);
        }
        h->gc_S = S; h->gc_threads = threads; h->gc_grid = grid; h->gc_smem = smem;
        return 1;
    }
    return 0;
}

static int grow_persistent(mcsce_handle h, const mcsce_grow_opts *o, void *d_workspace, size_t workspace_bytes, float *d_coords,
                           double *d_energy, uint8_t *d_ok, cudaStream_t st, bool any_plan) {
    const int C = o->n_conf;
    const int rc = gc_make_plan(h, C);
    if (rc <= 0) return rc;
    if (!any_plan && (long long)C * h->gc_S > h->sm_count) return 0;        // more than one round of clusters: the graph wins
    const SysDev &d = h->d;
    ClusterArgs a{};
    a.s = d; a.cp = CallParams{o->seed, o->conf_id0, o->beta, o->noise_deg};
    a.n_conf = C; a.cluster = h->gc_S; a.dedup_ok = o->no_dedup ? 0 : 1; a.max_trials = std::max(1, h->max_trials);
    a.env_pairs = h->env_pairs; a.pstart = h->d_pstart;
    a.unit_overhead = 0.10f;
    if (const char *e = getenv("MCSCE_B200_GC_UNIT_OVERHEAD")) a.unit_overhead = (float)atof(e);      // tuning knob
    if (h->trial_atoms_total >= 0x7fffffffLL) return 0;
    a.trial_stride = std::max<long long>(1, h->trial_atoms_total); a.aoff = h->d_aoff;
    const size_t need = gc_workspace_bytes(h, C);
    if (need > workspace_bytes || need > (size_t)GC_MAX_WORKSPACE) return 0;
    const size_t trial_bytes = align_up((size_t)C * (size_t)a.trial_stride * sizeof(float4));
    a.trial = (float4 *)d_workspace; a.sp = (double2 *)((char *)d_workspace + trial_bytes);
    a.chis_in = o->d_chis; a.ori_in = o->d_chis ? o->d_ori : nullptr; a.uniform = o->d_uniform;
    a.chis_out = o->d_chis_out; a.trial_energy = o->d_trial_energy; a.selected = o->d_selected;
    a.coords = d_coords; a.energy = d_energy; a.ok = d_ok; a.e_input = h->d_input_energy; a.pair_counter = h->d_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)h->gc_grid); cfg.blockDim = dim3((unsigned)h->gc_threads); cfg.dynamicSmemBytes = h->gc_smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (h->gc_S > 1) {
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)h->gc_S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    CK(cudaFuncSetAttribute(grow_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gc_smem));
    static long long *d_dbg = nullptr;
    const bool debug = getenv("MCSCE_B200_GC_DEBUG") != nullptr;
    if (debug) {
        if (!d_dbg) CK(cudaMalloc((void **)&d_dbg, 16 * sizeof(long long)));
        CK(cudaMemsetAsync(d_dbg, 0, 16 * sizeof(long long), st));
        a.debug = d_dbg;
    }
    CK(cudaLaunchKernelEx(&cfg, grow_cluster_kernel, a));
    h->launches += 1; h->persistent_launches += 1;
    if (debug) {
        long long t[16];
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(t, d_dbg, sizeof(t), cudaMemcpyDeviceToHost));
        static const char *nm[10] = {"init", "K1", "K2s", "prep", "generic", "reduce", "barrier", "select-min", "weights+scan", "append"};
        fprintf(stderr, "[mcsce_b200] persistent growth C=%d cluster=%d threads=%d smem=%zu, kclocks of conformer 0:", C, h->gc_S, h->gc_threads, h->gc_smem);
        for (int k = 0; k < 10; ++k) fprintf(stderr, " %s %.1f", nm[k], t[k] * 1e-3);
        fprintf(stderr, "\n");
    }
    return 1;
}

extern "C" int mcsce_grow(mcsce_handle h, const mcsce_grow_opts *o, void *d_workspace, size_t workspace_bytes,
                          float *d_coords, double *d_energy, uint8_t *d_ok, void *stream) {
    if (!h || !o) return fail("null argument");
    if (!h->has_bb) return fail("set_system/set_backbone first");
    if (o->n_conf <= 0) return fail("n_conf must be positive");
    if (o->n_conf > 65535) return fail("n_conf too large: at most 65535 conformers per call (conformers are the grid's y dimension); split the call");
    CK(cudaSetDevice(h->device));
    const SysDev &d = h->d;
    if (workspace_bytes < mcsce_workspace_bytes(h, o->n_conf)) return fail("workspace too small: call mcsce_workspace_bytes");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = o->n_conf;
    CK(cudaStreamWaitEvent(st, h->bb_ready, 0));       // set_backbone's copy and table kernels run on the default stream
    if (prepare_lane(h, h->lane[0], false)) return -1;

    // Launch-bound runs (little work per residue: few conformers and/or a small protein) replay a captured CUDA graph of
    // the whole growth instead of ~4 launches + 7 event calls per residue; the graph is cached per (tables, n_conf,
    // buffers, options).  Large runs enqueue directly and may re-derive the environment split while conformers die.
    long long grow_steps = 0, pairs_per_step = 0;
    const bool small = launch_bound(h, C, &grow_steps, &pairs_per_step);
    // (the coordinate-based term sizes its buffers at the first step of a growth: not capturable)
    // o->no_graph: 0 = automatic, 1 = the direct per-kernel path, 2 = a replayed graph wherever the automatic choice is a
    // launch-bound path, 3 = the persistent cluster kernel whatever the size (if the environment fits shared memory)
    const bool plain = !o->hpmf && !h->profiling && grow_steps > 0;
    const int persistent = persistent_knob();
    if (plain && (o->no_graph == 3 || (o->no_graph == 0 && ((small && persistent >= 1) || persistent >= 2)))) {
        // On its own the persistent kernel is picked where it was measured faster than the replayed graph (B200, ms per
        // ensemble, persistent / graph).  76 residues: 1 conformer 0.94 / 1.50, 4: 0.94 / 1.55, 8: 1.20 / 1.59, 16: 1.67 / 1.73,
        // 32: 1.68 / 2.07, 64: 2.56 / 2.87, 128: 4.33 / 4.52, 200 (two rounds of clusters): 8.83 / 6.18.  150 residues: 1: 1.83 /
        // 2.06, 4: 2.79 / 3.29, 8: 4.15 / 3.65.  300 residues: 1: 7.27 / 6.21, 4: 7.30 / 7.19 - with that much work per residue the
        // per-kernel path's environment split over all SMs beats 16 CTAs per conformer.  Hence: small proteins (<= 8e5 pair
        // evaluations per residue and conformer; <= 2e6 for up to 4 conformers) whose clusters are all resident at once.
        const bool any = (o->no_graph == 3 || persistent >= 2);
        const bool fits_rule = pairs_per_step <= 800000LL || (pairs_per_step <= 2000000LL && C <= 4);
        const int rc = (any || fits_rule) ? grow_persistent(h, o, d_workspace, workspace_bytes, d_coords, d_energy, d_ok, st, any) : 0;
        if (rc != 0) return rc < 0 ? -1 : 0;
    }
    const bool use_graph = (o->no_graph == 0 || o->no_graph == 2) && plain && small;
    if (use_graph) {
        Workspace w = carve(h, C, h->max_trials, h->max_trial_atoms, d_workspace);
        std::vector<long long> key = {h->table_version, C, (long long)(size_t)d_workspace, (long long)workspace_bytes,
                                      (long long)(size_t)o->d_chis, (long long)(size_t)o->d_uniform,
                                      (long long)(size_t)o->d_trial_energy, (long long)(size_t)o->d_chis_out,
                                      (long long)(size_t)o->d_selected, o->env_split, o->no_dedup, o->no_screen,
                                      (long long)(size_t)o->d_ori};
        if (!h->graph_exec || key != h->graph_key) {
            if (!h->cap_stream) CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
            const long long before = h->launches;
            h->capturing = true;
            CK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed));
            const int rc = enqueue_growth(h, o, w, h->lane[0], h->cap_stream, h->lane[0].aux, false);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
            h->capturing = false;
            h->graph_launches = h->launches - before;
            h->launches = before;
            if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
            if (e != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            // a new backbone of the same sequence gives the same graph topology with other launch sizes and table
            // pointers: update the instantiated graph in place (much cheaper than instantiating), else start over
            bool updated = false;
            if (h->graph_exec) {
                cudaGraphExecUpdateResultInfo info;
                updated = cudaGraphExecUpdate(h->graph_exec, graph, &info) == cudaSuccess;
                if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
            }
            if (!updated) e = cudaGraphInstantiate(&h->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { h->graph_exec = nullptr; return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
            h->graph_key = key;
        }
        CallParams cp{o->seed, o->conf_id0, o->beta, o->noise_deg};
        set_call_params_kernel<<<1, 1, 0, st>>>(h->lane[0].d_call, cp);
        h->launches += 1;
        CK(cudaGraphLaunch(h->graph_exec, st));
        h->launches += h->graph_launches;
        return finish_growth(h, C, w, d_coords, d_energy, d_ok, st);
    }

    const int G = h->profiling ? 1 : lane_count(h, C);
    if (G == 1) {
        Workspace w = carve(h, C, h->max_trials, h->max_trial_atoms, d_workspace);
        // while the per-class event timing is on, everything goes to one stream so that each kernel is timed alone
        if (enqueue_growth(h, o, w, h->lane[0], st, h->profiling ? st : h->lane[0].aux, true)) return -1;
        return finish_growth(h, C, w, d_coords, d_energy, d_ok, st, o->hpmf ? h->lane[0].hp_ovf : nullptr);
    }
    // G groups of conformers, each a complete growth on its own pair of streams; group 0 on the caller's stream.  The groups
    // are enqueued interleaved, residue by residue, and share one environment-split rule (that of the whole call).
    const long long chi_stride = (long long)d.n_trials_total * MAX_CHI;
    size_t ws_off = 0;
    for (int g = 1; g < G; ++g) {                               // fork: the other groups start after whatever precedes the call on `st`
        auto &ln = h->lane[g];
        if (prepare_lane(h, ln, true)) return -1;
        CK(cudaEventRecord(ln.ev0, st));
        CK(cudaStreamWaitEvent(ln.main, ln.ev0, 0));
    }
    std::vector<GrowthEnqueue> ge;
    std::vector<int> c0s, cgs;
    for (int g = 0; g < G; ++g) {
        auto &ln = h->lane[g];
        const int c0 = (int)((long long)C * g / G), cg = (int)((long long)C * (g + 1) / G) - c0;
        Workspace w = carve(h, cg, h->max_trials, h->max_trial_atoms, (char *)d_workspace + ws_off);
        ws_off += w.total;
        mcsce_grow_opts og = *o;
        og.n_conf = cg; og.conf_id0 = o->conf_id0 + c0;
        if (o->d_chis) og.d_chis = o->d_chis + (long long)c0 * chi_stride;
        if (o->d_chis_out) og.d_chis_out = o->d_chis_out + (long long)c0 * chi_stride;
        if (o->d_ori) og.d_ori = o->d_ori + (long long)c0 * (chi_stride / MAX_CHI) * MAX_ROT;
        if (o->d_uniform) og.d_uniform = o->d_uniform + (long long)c0 * d.n_res;
        if (o->d_selected) og.d_selected = o->d_selected + (long long)c0 * d.n_res;
        if (o->d_trial_energy) og.d_trial_energy = o->d_trial_energy + (long long)c0 * d.n_trials_total;
        ge.push_back(GrowthEnqueue{h, og, w, &ln, g == 0 ? st : ln.main, ln.aux, true, G});
        c0s.push_back(c0); cgs.push_back(cg);
    }
    for (auto &g : ge) if (g.begin()) return -1;
    for (int i = 0; i < d.n_res; ++i)
        for (auto &g : ge) if (g.residue(i)) return -1;
    CK(cudaGetLastError());
    for (int g = G - 1; g >= 0; --g) {
        if (finish_growth(h, cgs[g], ge[g].w, d_coords + (size_t)c0s[g] * d.n_atoms * 3, d_energy + c0s[g], d_ok + c0s[g], ge[g].st,
                          o->hpmf ? h->lane[g].hp_ovf : nullptr)) return -1;
        if (g > 0) { CK(cudaEventRecord(h->lane[g].done, ge[g].st)); CK(cudaStreamWaitEvent(st, h->lane[g].done, 0)); }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Mode reductions of the callers (core/side_chain_builder.py:245-249 "simple": the first successful conformer;
// :280-281 "exhaustive": the successful conformer of lowest energy) on the device: the index, and the picked conformer's
// coordinates, so that only one conformer travels to the host instead of the ensemble.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pick_kernel(int n_conf, int n_atoms, const double *energy, const unsigned char *ok,
                                                   const float *coords, int first_valid, int *index, float *picked) {
    __shared__ double s_e[256];
    __shared__ int s_i[256];
    const int tid = threadIdx.x;
    double best = CUDART_INF; int bi = 0x7fffffff;
    for (int i = tid; i < n_conf; i += 256) {
        if (ok[i] != 1) continue;
        const double e = first_valid ? 0.0 : energy[i];
        if (bi == 0x7fffffff || e < best || (e == best && i < bi)) { best = e; bi = i; }   // ties -> the lowest index, like np.argmin / np.argmax
    }
    s_e[tid] = best; s_i[tid] = bi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) {
            const double e2 = s_e[tid + w]; const int i2 = s_i[tid + w];
            if (i2 != 0x7fffffff && (s_i[tid] == 0x7fffffff || e2 < s_e[tid] || (e2 == s_e[tid] && i2 < s_i[tid]))) { s_e[tid] = e2; s_i[tid] = i2; }
        }
        __syncthreads();
    }
    const int sel = s_i[0] == 0x7fffffff ? -1 : s_i[0];
    if (tid == 0) index[0] = sel;
    if (sel >= 0 && picked)
        for (int k = tid; k < n_atoms * 3; k += 256) picked[k] = coords[(long long)sel * n_atoms * 3 + k];
}

extern "C" int mcsce_pick(mcsce_handle h, int n_conf, const double *d_energy, const uint8_t *d_ok, const float *d_coords,
                          int first_valid, int32_t *d_index, float *d_picked_coords, void *stream) {
    if (!h || !h->has_sys) return fail("set_system first");
    if (n_conf <= 0 || !d_energy || !d_ok || !d_index) return fail("bad arguments");
    if (d_picked_coords && !d_coords) return fail("d_coords is needed to copy the picked conformer");
    CK(cudaSetDevice(h->device));
    pick_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(n_conf, h->d.n_atoms, d_energy, d_ok, d_coords, first_valid, d_index, d_picked_coords);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int mcsce_grow_host(mcsce_handle h, const mcsce_grow_opts *o, void *d_workspace, size_t workspace_bytes,
                               float *h_coords, double *h_energy, uint8_t *h_ok, void *stream) {
    if (!h || !o) return fail("null argument");
    if (!h->has_bb) return fail("set_system/set_backbone first");
    const size_t base = mcsce_workspace_bytes(h, o->n_conf);
    const size_t nc = align_up((size_t)o->n_conf * h->d.n_atoms * 3 * sizeof(float));
    const size_t ne = align_up((size_t)o->n_conf * sizeof(double));
    const size_t nk = align_up((size_t)o->n_conf);
    if (workspace_bytes < base + nc + ne + nk) return fail("workspace too small for mcsce_grow_host");
    char *p = (char *)d_workspace + base;
    float *dc = (float *)p; double *de = (double *)(p + nc); uint8_t *dk = (uint8_t *)(p + nc + ne);
    if (mcsce_grow(h, o, d_workspace, base, dc, de, dk, stream)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h_coords, dc, (size_t)o->n_conf * h->d.n_atoms * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_energy, de, (size_t)o->n_conf * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_ok, dk, (size_t)o->n_conf, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (o->hpmf) {
        int n_over = 0;
        for (int i = 0; i < o->n_conf; ++i) n_over += h_ok[i] == 2;
        if (n_over) return fail("hydrophobic PMF: a neighbour-list capacity was exceeded for " + std::to_string(n_over) +
                                " conformers (ok = 2); the other conformers are complete");
    }
    return 0;
}

extern "C" int mcsce_rotate_template(int device, int n_atoms, const float *h_tmpl_xyz, int n_rot, const int32_t *h_axis,
                                     const uint32_t *h_move, const double *h_ori, int n_rows, const double *h_chis, float *h_out) {
    // replaces rotate_sidechain (libs/libscbuild.py:56-206): host buffers in/out, the rotations run on the device
    if (n_atoms <= 0 || n_atoms > MAX_TMPL || n_rot < 0 || n_rot > MAX_ROT || n_rows <= 0) return fail("bad template arguments");
    CK(cudaSetDevice(device));
    TmplDev tm; memset(&tm, 0, sizeof(tm));
    tm.n_atoms = n_atoms; tm.n_chi = n_rot;
    for (int a = 0; a < n_atoms; ++a) for (int k = 0; k < 3; ++k) tm.xyz[a][k] = h_tmpl_xyz[a * 3 + k];
    for (int k = 0; k < n_rot; ++k) { tm.axis_from[k] = h_axis[2 * k]; tm.axis_to[k] = h_axis[2 * k + 1]; tm.move[k] = h_move[k]; tm.ori[k] = h_ori[k]; }
    double *d_chis = nullptr; float *d_out = nullptr;
    CK(cudaMalloc((void **)&d_chis, sizeof(double) * (size_t)n_rows * MAX_CHI));
    CK(cudaMalloc((void **)&d_out, sizeof(float) * (size_t)n_rows * n_atoms * 3));
    CK(cudaMemcpy(d_chis, h_chis, sizeof(double) * (size_t)n_rows * MAX_CHI, cudaMemcpyHostToDevice));
    rotate_template_kernel<<<(n_rows + 3) / 4, 128>>>(tm, n_rot, n_rows, d_chis, d_out);
    cudaError_t e = cudaMemcpy(h_out, d_out, sizeof(float) * (size_t)n_rows * n_atoms * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_chis); cudaFree(d_out);
    if (e != cudaSuccess) return fail(std::string("rotate_template: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int mcsce_place_template(int device, int n_atoms, const float *h_xyz, const double *h_q1, const double *h_q2,
                                    const float *h_ca, double *h_out) {
    // replaces the coordinate part of place_sidechain_template (libs/libcalc.py:465-482)
    if (n_atoms <= 0) return fail("bad template arguments");
    CK(cudaSetDevice(device));
    float *d_in = nullptr; double *d_out = nullptr;
    CK(cudaMalloc((void **)&d_in, sizeof(float) * (size_t)n_atoms * 3));
    CK(cudaMalloc((void **)&d_out, sizeof(double) * (size_t)n_atoms * 3));
    CK(cudaMemcpy(d_in, h_xyz, sizeof(float) * (size_t)n_atoms * 3, cudaMemcpyHostToDevice));
    place_template_kernel<<<(n_atoms + 63) / 64, 64>>>(n_atoms, d_in, h_q1[0], h_q1[1], h_q1[2], h_q1[3], h_q2[0], h_q2[1],
                                                       h_q2[2], h_q2[3], h_ca[0], h_ca[1], h_ca[2], d_out);
    cudaError_t e = cudaMemcpy(h_out, d_out, sizeof(double) * (size_t)n_atoms * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(std::string("place_template: ") + cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Peer-memory result buffer: the multi-GPU "gather" without a collective.  Rank 0 owns one buffer for the whole ensemble;
// every other rank maps it (CUDA IPC over NVLink/NVSwitch) and passes its slice as d_coords / d_energy / d_ok of mcsce_grow,
// so the result kernels (gather_kernel, finish_kernel) store straight into rank 0's HBM - the transfer is the kernel's own
// store stream, there is no staging copy and no second pass.  Replaces what multiprocessing.Pool.imap_unordered ships back to
// the parent in the reference (core/side_chain_builder.py:328-347).
// ------------------------------------------------------------------------------------------------
extern "C" int mcsce_peer_alloc(int device, size_t bytes, void **d_ptr, unsigned char *handle64) {
    if (!d_ptr || !handle64 || bytes == 0) return fail("mcsce_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(device));
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(handle64, &hd, 64);
    *d_ptr = p;
    return 0;
}

extern "C" int mcsce_peer_open(int device, const unsigned char *handle64, void **d_ptr) {
    if (!d_ptr || !handle64) return fail("mcsce_peer_open: bad arguments");
    CK(cudaSetDevice(device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return 0;
}

extern "C" int mcsce_peer_close(int device, void *d_ptr) {
    CK(cudaSetDevice(device));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

extern "C" int mcsce_peer_free(int device, void *d_ptr) {
    CK(cudaSetDevice(device));
    CK(cudaFree(d_ptr));
    return 0;
}

extern "C" int mcsce_profile(mcsce_handle h, int enable, double *ms_by_class, int64_t *launches_by_class) {
    // enable=1: start recording event pairs around every K1/K2/K2s/K3 launch.
    // enable=0: synchronise, return summed milliseconds and launch counts per class (4 each), stop.
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    if (enable) {
        for (auto &v : h->ev) { for (auto e : v) cudaEventDestroy(e); v.clear(); }
        h->profiling = true;
        return 0;
    }
    h->profiling = false;
    CK(cudaDeviceSynchronize());
    for (int c = 0; c < 5; ++c) {
        double ms = 0.0;
        for (size_t i = 0; i + 1 < h->ev[c].size(); i += 2) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, h->ev[c][i], h->ev[c][i + 1]));
            ms += t;
        }
        if (ms_by_class) ms_by_class[c] = ms;
        if (launches_by_class) launches_by_class[c] = (int64_t)(h->ev[c].size() / 2);
        for (auto e : h->ev[c]) cudaEventDestroy(e);
        h->ev[c].clear();
    }
    return 0;
}

extern "C" int mcsce_counters(mcsce_handle h, int64_t *launches, int64_t *pairs, int64_t *pairs_computed, int reset) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    unsigned long long v[2] = {0, 0};
    CK(cudaMemcpy(v, h->d_pairs, sizeof(v), cudaMemcpyDeviceToHost));
    if (launches) *launches = h->launches;
    if (pairs) *pairs = (int64_t)v[0];
    if (pairs_computed) *pairs_computed = (int64_t)v[1];
    if (reset) { h->launches = 0; CK(cudaMemset(h->d_pairs, 0, sizeof(v))); }
    return 0;
}
