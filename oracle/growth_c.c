/*
 * ORACLE - test infrastructure only (C restatement of the MCSCE growth hot path).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may load
 * this library; the product (mcsce_b200) never does.  Built by mcsce_b200/build.py:build_oracle_c
 * with gcc -O2 -ffp-contract=off -fopenmp.
 *
 * It follows the reference function by function, scalar C, one conformer per OpenMP thread (the
 * reference runs one conformer per multiprocessing worker, core/side_chain_builder.py:328-347):
 *   set_chis()        <- rotate_sidechain / rotate_tor          libs/libscbuild.py:17-206
 *   dihedral_f32()    <- calc_torsion_angles on 4 atoms         libs/libcalc.py:271-304
 *   rotate_q()        <- rotate_coordinates_Q + Hamilton product libs/libcalc.py:358-417
 *   place()           <- place_sidechain_template (application of the two rotations) libs/libcalc.py:465-482
 *   trial_energy()    <- calc_new_vs_old_dists + clash filter + LJ + Coulomb
 *                        libs/libcalc.py:47-78, libs/libenergy.py:680-689, 717-724, 786-810
 *   choose()          <- choose_random                          core/side_chain_builder.py:33-40
 *   grow_one()        <- create_side_chain_structure            core/side_chain_builder.py:109-205
 * Per-pair parameters come in the factored form the host tables use (atom classes + the exact
 * special-pair list), which tests/test_system_tables.py proves equal to the reference's per-pair
 * arrays.  PARITY PIN: validated against oracle/ref_numpy.py (itself pinned to the running
 * reference by tests/golden) in tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_TMPL 32
#define MAX_SC 28
#define MAX_ROT 5
#define MAX_CHI 6
#define MAX_SP_ENV 16
#define PI 3.14159265358979323846

typedef struct {
    int n_atoms, n_input, n_res, n_cls, n_types;
    const int *cls_of_atom; const double *charge;
    const double *cls_s2, *cls_e4, *cls_thr;
    const int *tmpl_n_atoms; const float *tmpl_xyz; const int *tmpl_sc_pos;
    const int *tmpl_dih, *tmpl_axis; const unsigned *tmpl_move;
    const int *step_kind, *step_type, *step_n_sc, *step_sc_start, *step_n_sp_env, *step_sp_env, *step_sp_off, *step_n_sp;
    const int *sp_k, *sp_partner; const double *sp_coef;
    int n_bb_pairs; const int *bb_i, *bb_j; const double *bb_coef; const int *resnum;
    const int *step_n_trials, *step_trial_off, *step_n_chi_lib, *step_n_chi;
    const double *chi_mean, *chi_sigma, *prob;
    const double *place_q1, *place_q2;   /* n_res*4 each */
    const float *place_ca;               /* n_res*3 */
    const float *input_xyz;
    int n_trials_total;
} oracle_system;

/* ---- float32 helpers (unfused: build with -ffp-contract=off) ---- */
static void cross32(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static float norm32(const float *v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static float dot32(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* np.linalg.norm of a float32 3-vector as NumPy + OpenBLAS evaluate it (libscbuild.py:115,169,185,199 call it on the
 * rotation axis): sqrt(x.dot(x)) with sdot's scalar tail - float32 products added up in a double, the sum rounded to
 * float32, float32 sqrt.  Checked against np.linalg.norm on 20000 random vectors: identical. */
static float norm32_blas(const float *v) {
    const float xx = v[0] * v[0], yy = v[1] * v[1], zz = v[2] * v[2];
    const double acc = ((double)xx + (double)yy) + (double)zz;
    return sqrtf((float)acc);
}

static float dihedral_f32(const float *p0, const float *p1, const float *p2, const float *p3) {
    float q0[3], q1[3], q2[3], c0[3], c1[3], u3[3], u2[3];
    for (int k = 0; k < 3; ++k) { q0[k] = p1[k] - p0[k]; q1[k] = p2[k] - p1[k]; q2[k] = p3[k] - p2[k]; }
    cross32(q0, q1, c0); cross32(q1, q2, c1);
    float n0 = norm32(c0), n1 = norm32(c1), nq = norm32(q1);
    for (int k = 0; k < 3; ++k) { c0[k] /= n0; c1[k] /= n1; u3[k] = q1[k] / nq; }
    cross32(u3, c1, u2);
    return -atan2f(dot32(c0, u2), dot32(c0, c1));
}

static void rotate_q(double b1, double b2, double b3, double b4, double x, double y, double z, double *o) {
    double c1 = (b1 * 0.0) - (b2 * x) - (b3 * y) - (b4 * z);
    double c2 = (b1 * x) + (b2 * 0.0) + (b3 * z) - (b4 * y);
    double c3 = (b1 * y) - (b2 * z) + (b3 * 0.0) + (b4 * x);
    double c4 = (b1 * z) + (b2 * y) - (b3 * x) + (b4 * 0.0);
    double n2 = -b2, n3 = -b3, n4 = -b4;
    o[0] = (c1 * n2) + (c2 * b1) + (c3 * n4) - (c4 * n3);
    o[1] = (c1 * n3) - (c2 * n4) + (c3 * b1) + (c4 * n2);
    o[2] = (c1 * n4) + (c2 * n3) - (c3 * n2) + (c4 * b1);
}

/* ori_in: measured dihedrals to use instead of re-measuring (the reference's recorded float32 values; NaN = measure);
 * ori_out: the values used.  Either may be NULL. */
static void set_chis(const oracle_system *s, int type, int n_rot, const double *tors_deg, float xyz[MAX_TMPL][3],
                     const double *ori_in, double *ori_out) {
    const int n = s->tmpl_n_atoms[type];
    memcpy(xyz, s->tmpl_xyz + (size_t)type * MAX_TMPL * 3, sizeof(float) * MAX_TMPL * 3);
    for (int k = 0; k < n_rot; ++k) {
        const int *dih = s->tmpl_dih + ((size_t)type * MAX_ROT + k) * 4;
        const int af = s->tmpl_axis[((size_t)type * MAX_ROT + k) * 2], at = s->tmpl_axis[((size_t)type * MAX_ROT + k) * 2 + 1];
        const unsigned mv = s->tmpl_move[(size_t)type * MAX_ROT + k];
        double measured = (double)dihedral_f32(xyz[dih[0]], xyz[dih[1]], xyz[dih[2]], xyz[dih[3]]);
        if (ori_in && ori_in[k] == ori_in[k]) measured = ori_in[k];
        if (ori_out) ori_out[k] = measured;
        float piv[3], ax[3];
        for (int c = 0; c < 3; ++c) { piv[c] = xyz[at][c]; ax[c] = xyz[at][c] - xyz[af][c]; }
        const float nrm = norm32_blas(ax);
        for (int c = 0; c < 3; ++c) ax[c] = ax[c] / nrm;
        double rot = measured - tors_deg[k] * PI / 180;
        if (rot > PI) rot -= 2 * PI; else if (rot < -PI) rot += 2 * PI;
        const double sn = sin(rot / 2), cs = cos(rot / 2);
        const double b2 = sn * (double)(-ax[0]), b3 = sn * (double)(-ax[1]), b4 = sn * (double)(-ax[2]);
        for (int a = 0; a < n; ++a) {
            if (!((mv >> a) & 1u)) continue;
            float d[3]; double o[3];
            for (int c = 0; c < 3; ++c) d[c] = xyz[a][c] - piv[c];
            rotate_q(cs, b2, b3, b4, (double)d[0], (double)d[1], (double)d[2], o);
            for (int c = 0; c < 3; ++c) xyz[a][c] = (float)(o[c] + (double)piv[c]);
        }
    }
}

/* Rigid placement: the two quaternions depend only on the backbone and on template atoms 0 and 2, so
 * the Python side derives them once per residue with the NumPy restatement of libcalc.py:449-480
 * (float32 BLAS norm/dot like the jitted reference - plain C loops round differently and acos
 * amplifies that) and this function applies them: first rotation stored to float32 (libcalc.py:465),
 * second in float64, + CA, stored to float32 (side_chain_builder.py:181). */
static void place(const double *q1, const double *q2, const float *ca, int n, float xyz[MAX_TMPL][3], float out[MAX_TMPL][3]) {
    double o[3];
    for (int a = 0; a < n; ++a) {
        rotate_q(q1[0], q1[1], q1[2], q1[3], (double)xyz[a][0], (double)xyz[a][1], (double)xyz[a][2], o);
        const float r0 = (float)o[0], r1 = (float)o[1], r2 = (float)o[2];
        rotate_q(q2[0], q2[1], q2[2], q2[3], (double)r0, (double)r1, (double)r2, o);
        for (int c = 0; c < 3; ++c) out[a][c] = (float)(o[c] + (double)ca[c]);
    }
}

static inline float dist2(const float *a, const float *b) {
    const float x = b[0] - a[0], y = b[1] - a[1], z = b[2] - a[2];
    return x * x + y * y + z * z;
}

static inline double pair_e(double d2, double d, double A, double B, double Q) {
    const double d6 = d2 * d2 * d2;
    double e = 0.0;
    if (A != 0.0 || B != 0.0) e = A / (d6 * d6) - B / d6;
    if (Q != 0.0) e += Q / d;
    return e;
}

/* energy of one trial side chain `tr` (n_sc x 3) against env `xyz` (growth order, first n_old atoms);
 * returns +inf on clash.  Two passes like the reference: distances + clash filter, then energies. */
static double trial_energy(const oracle_system *s, int step, const float *tr, const float *xyz, int n_old,
                           const unsigned char *is_special, const double *clsA, const double *clsB) {
    const int n_sc = s->step_n_sc[step], sc0 = s->step_sc_start[step], nc = s->n_cls;
    const int off = s->step_sp_off[step], nsp = s->step_n_sp[step];
    const int *spenv = s->step_sp_env + (size_t)step * MAX_SP_ENV;
    for (int p = 0; p < nsp; ++p) {
        const int k = s->sp_k[off + p], pa = s->sp_partner[off + p];
        const float *b = pa >= 0 ? xyz + (size_t)spenv[pa] * 3 : tr + (size_t)(-pa - 1) * 3;
        const double d = sqrt((double)dist2(tr + (size_t)k * 3, b));
        if (d < s->sp_coef[(size_t)(off + p) * 4 + 3]) return INFINITY;
    }
    for (int k = 0; k < n_sc; ++k) {
        const int ci = s->cls_of_atom[sc0 + k];
        const double *thr = s->cls_thr + (size_t)ci * nc;
        for (int j = 0; j < n_old; ++j) {
            if (is_special[j]) continue;
            const double d = sqrt((double)dist2(tr + (size_t)k * 3, xyz + (size_t)j * 3));
            if (d < thr[s->cls_of_atom[j]]) return INFINITY;
        }
    }
    double e = 0.0;
    for (int p = 0; p < nsp; ++p) {
        const int k = s->sp_k[off + p], pa = s->sp_partner[off + p];
        const float *b = pa >= 0 ? xyz + (size_t)spenv[pa] * 3 : tr + (size_t)(-pa - 1) * 3;
        const double d2 = (double)dist2(tr + (size_t)k * 3, b);
        const double *cf = s->sp_coef + (size_t)(off + p) * 4;
        e += pair_e(d2, sqrt(d2), cf[0], cf[1], cf[2]);
    }
    for (int k = 0; k < n_sc; ++k) {
        const int ci = s->cls_of_atom[sc0 + k];
        const double qi = s->charge[sc0 + k] * 0.25 * 1389.35;
        for (int j = 0; j < n_old; ++j) {
            if (is_special[j]) continue;
            const int cj = s->cls_of_atom[j];
            const double d2 = (double)dist2(tr + (size_t)k * 3, xyz + (size_t)j * 3);
            e += pair_e(d2, sqrt(d2), clsA[ci * nc + cj], clsB[ci * nc + cj], qi * s->charge[j]);
        }
    }
    return e;
}

static int choose(const double *w, int n, double u) {
    double total = 0.0;
    for (int i = 0; i < n; ++i) total += w[i];
    const double pointer = u * total;
    double run = 0.0;
    for (int i = 0; i < n; ++i) { run += w[i]; if (run > pointer) return i; }
    return n - 1;
}

/* splitmix64 + Box-Muller: CPU-baseline runs only (parity runs feed chis and uniforms) */
static uint64_t sm64(uint64_t *st) { uint64_t z = (*st += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static double u01(uint64_t *st) { return (double)(sm64(st) >> 11) * (1.0 / 9007199254740992.0); }
static double gauss(uint64_t *st) { double u1 = 1.0 - u01(st), u2 = u01(st); return sqrt(-2.0 * log(u1)) * cos(2 * PI * u2); }

double oracle_input_energy(const oracle_system *s) {
    const int nc = s->n_cls;
    double e = 0.0;
    for (int p = 0; p < s->n_bb_pairs; ++p) {
        const double d2 = (double)dist2(s->input_xyz + (size_t)s->bb_i[p] * 3, s->input_xyz + (size_t)s->bb_j[p] * 3);
        const double *cf = s->bb_coef + (size_t)p * 4;
        if (sqrt(d2) < cf[3]) return INFINITY;
        e += pair_e(d2, sqrt(d2), cf[0], cf[1], cf[2]);
    }
    for (int i = 0; i < s->n_input; ++i)
        for (int j = i + 1; j < s->n_input; ++j) {
            const int dr = s->resnum[j] - s->resnum[i];
            if (dr >= -1 && dr <= 1) continue;
            const int ci = s->cls_of_atom[i], cj = s->cls_of_atom[j];
            const double d2 = (double)dist2(s->input_xyz + (size_t)i * 3, s->input_xyz + (size_t)j * 3);
            if (sqrt(d2) < s->cls_thr[ci * nc + cj]) return INFINITY;
            const double s2 = s->cls_s2[ci * nc + cj], e4 = s->cls_e4[ci * nc + cj], s6 = s2 * s2 * s2;
            e += pair_e(d2, sqrt(d2), e4 * s6 * s6, e4 * s6, s->charge[i] * s->charge[j] * 0.25 * 1389.35);
        }
    return e;
}

static void grow_one(const oracle_system *s, int conf, double beta, double noise_deg, uint64_t seed, const double *chis_in,
                     const double *uniforms, double e_input, const double *clsA, const double *clsB, float *coords,
                     double *energy, unsigned char *ok, double *trial_energy_out, int *selected, float *trial_xyz_out,
                     const double *ori_in, double *ori_out) {
    float *xyz = coords + (size_t)conf * s->n_atoms * 3;
    memcpy(xyz, s->input_xyz, sizeof(float) * 3 * s->n_input);
    uint64_t rng = seed * 0x9E3779B97F4A7C15ull + (uint64_t)conf * 0xD1B54A32D192ED03ull + 1;
    double e_acc = e_input;
    unsigned char *is_special = (unsigned char *)calloc(s->n_atoms, 1);
    int max_t = 1;
    for (int i = 0; i < s->n_res; ++i) if (s->step_n_trials[i] > max_t) max_t = s->step_n_trials[i];
    float *trial = (float *)malloc(sizeof(float) * (size_t)max_t * MAX_SC * 3);
    double *E = (double *)malloc(sizeof(double) * max_t), *w = (double *)malloc(sizeof(double) * max_t);
    int alive = 1;
    for (int i = 0; i < s->n_res && alive; ++i) {
        if (selected) selected[(size_t)conf * s->n_res + i] = -1;
        const int kind = s->step_kind[i];
        if (kind == 0) continue;
        const int type = s->step_type[i], n_sc = s->step_n_sc[i], sc0 = s->step_sc_start[i], nt = s->tmpl_n_atoms[type];
        const double *q1 = s->place_q1 + (size_t)i * 4, *q2 = s->place_q2 + (size_t)i * 4;
        const float *ca = s->place_ca + (size_t)i * 3;
        float tx[MAX_TMPL][3], placed[MAX_TMPL][3];
        if (kind == 1) {                       /* GLY/ALA: rigid placement, no energy (side_chain_builder.py:164-168) */
            set_chis(s, type, 0, NULL, tx, NULL, NULL);
            place(q1, q2, ca, nt, tx, placed);
            for (int k = 0; k < n_sc; ++k) memcpy(xyz + (size_t)(sc0 + k) * 3, placed[s->tmpl_sc_pos[(size_t)type * MAX_SC + k]], 12);
            continue;
        }
        const int T = s->step_n_trials[i], off = s->step_trial_off[i], nlib = s->step_n_chi_lib[i], nrot = s->step_n_chi[i];
        for (int q = 0; q < s->step_n_sp_env[i]; ++q) is_special[s->step_sp_env[(size_t)i * MAX_SP_ENV + q]] = 1;
        double emin = INFINITY;
        for (int t = 0; t < T; ++t) {
            double chi[MAX_CHI];
            for (int k = 0; k < MAX_CHI; ++k) {
                if (chis_in) chi[k] = chis_in[((size_t)conf * s->n_trials_total + off + t) * MAX_CHI + k];
                else if (k < nlib) {
                    double v = s->chi_mean[(size_t)(off + t) * MAX_CHI + k];
                    const double sg = s->chi_sigma[(size_t)(off + t) * MAX_CHI + k];
                    if (sg != 0.0) { v += sg * gauss(&rng); if (v > 180) v -= 360; if (v < -180) v += 360; }
                    chi[k] = v + noise_deg * gauss(&rng);
                } else chi[k] = 0.0;
            }
            {
                const size_t row = ((size_t)conf * s->n_trials_total + off + t) * MAX_ROT;
                set_chis(s, type, nrot, chi, tx, ori_in ? ori_in + row : NULL, ori_out ? ori_out + row : NULL);
            }
            place(q1, q2, ca, nt, tx, placed);
            float *tr = trial + (size_t)t * n_sc * 3;
            for (int k = 0; k < n_sc; ++k) memcpy(tr + (size_t)k * 3, placed[s->tmpl_sc_pos[(size_t)type * MAX_SC + k]], 12);
            E[t] = trial_energy(s, i, tr, xyz, sc0, is_special, clsA, clsB);
            if (trial_energy_out) trial_energy_out[(size_t)conf * s->n_trials_total + off + t] = E[t];
            if (trial_xyz_out && conf == 0) memcpy(trial_xyz_out + (size_t)(off + t) * MAX_SC * 3, tr, sizeof(float) * n_sc * 3);
            if (E[t] < emin) emin = E[t];
        }
        for (int q = 0; q < s->step_n_sp_env[i]; ++q) is_special[s->step_sp_env[(size_t)i * MAX_SP_ENV + q]] = 0;
        if (isinf(emin)) { alive = 0; break; }
        for (int t = 0; t < T; ++t) w[t] = s->prob[off + t] * exp(-beta * (E[t] - emin));
        const double u = uniforms ? uniforms[(size_t)conf * s->n_res + i] : u01(&rng);
        const int sel = choose(w, T, u);
        if (selected) selected[(size_t)conf * s->n_res + i] = sel;
        memcpy(xyz + (size_t)sc0 * 3, trial + (size_t)sel * n_sc * 3, sizeof(float) * n_sc * 3);
        e_acc += (E[sel] - emin) + emin;
    }
    ok[conf] = (unsigned char)alive;
    energy[conf] = alive ? e_acc : NAN;
    free(is_special); free(trial); free(E); free(w);
}

int oracle_grow(const oracle_system *s, int n_conf, double beta, double noise_deg, uint64_t seed, const double *chis_in,
                const double *uniforms, float *coords, double *energy, unsigned char *ok, double *trial_energy_out,
                int *selected, float *trial_xyz_out, int n_threads, const double *ori_in, double *ori_out) {
    const int nc = s->n_cls;
    double *clsA = (double *)malloc(sizeof(double) * nc * nc), *clsB = (double *)malloc(sizeof(double) * nc * nc);
    for (int i = 0; i < nc * nc; ++i) {
        const double s6 = s->cls_s2[i] * s->cls_s2[i] * s->cls_s2[i];
        clsA[i] = s->cls_e4[i] * s6 * s6; clsB[i] = s->cls_e4[i] * s6;
    }
    const double e_input = oracle_input_energy(s);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < n_conf; ++c)
        grow_one(s, c, beta, noise_deg, seed, chis_in, uniforms, e_input, clsA, clsB, coords, energy, ok,
                 trial_energy_out, selected, trial_xyz_out, ori_in, ori_out);
    free(clsA); free(clsB);
    return 0;
}

/* Total energy of a finished conformer re-scored level by level (input atoms alone, then every grown side chain
 * against everything placed before it) and the smallest ratio distance / clash distance over all scored pairs.
 * Size-independent check for full-size runs: must equal the energy the growth accumulated, and the ratio must
 * be >= 1 for a conformer reported as grown without clashes. */
static double rescore_impl(const oracle_system *s, const float *xyz, double *min_clash_ratio, double *level_energy);

double oracle_rescore(const oracle_system *s, const float *xyz, double *min_clash_ratio) {
    return rescore_impl(s, xyz, min_clash_ratio, NULL);
}

/* Same, and level_energy[i] = energy of residue i's side chain against everything placed before it (NaN for
 * residues that are not grown): the per-trial quantity of libenergy.py:786-810 for the trial that was selected. */
double oracle_rescore_levels(const oracle_system *s, const float *xyz, double *min_clash_ratio, double *level_energy) {
    return rescore_impl(s, xyz, min_clash_ratio, level_energy);
}

static double rescore_impl(const oracle_system *s, const float *xyz, double *min_clash_ratio, double *level_energy) {
    const int nc = s->n_cls;
    double *clsA = (double *)malloc(sizeof(double) * nc * nc), *clsB = (double *)malloc(sizeof(double) * nc * nc);
    for (int i = 0; i < nc * nc; ++i) {
        const double s6 = s->cls_s2[i] * s->cls_s2[i] * s->cls_s2[i];
        clsA[i] = s->cls_e4[i] * s6 * s6; clsB[i] = s->cls_e4[i] * s6;
    }
    oracle_system tmp = *s;
    tmp.input_xyz = xyz;
    double total = oracle_input_energy(&tmp);
    double ratio = INFINITY;
    unsigned char *is_special = (unsigned char *)calloc(s->n_atoms, 1);
    for (int i = 0; i < s->n_res; ++i) {
        if (level_energy) level_energy[i] = NAN;
        if (s->step_kind[i] != 2) continue;
        const int n_sc = s->step_n_sc[i], sc0 = s->step_sc_start[i];
        for (int q = 0; q < s->step_n_sp_env[i]; ++q) is_special[s->step_sp_env[(size_t)i * MAX_SP_ENV + q]] = 1;
        const double e_level = trial_energy(s, i, xyz + (size_t)sc0 * 3, xyz, sc0, is_special, clsA, clsB);
        if (level_energy) level_energy[i] = e_level;
        total += e_level;
        for (int k = 0; k < n_sc; ++k) {
            const int ci = s->cls_of_atom[sc0 + k];
            for (int j = 0; j < sc0; ++j) {
                if (is_special[j]) continue;
                const double thr = s->cls_thr[ci * nc + s->cls_of_atom[j]];
                if (thr > 0) { const double r = sqrt((double)dist2(xyz + (size_t)(sc0 + k) * 3, xyz + (size_t)j * 3)) / thr; if (r < ratio) ratio = r; }
            }
        }
        for (int p = 0; p < s->step_n_sp[i]; ++p) {
            const int off = s->step_sp_off[i];
            const int k = s->sp_k[off + p], pa = s->sp_partner[off + p];
            const double thr = s->sp_coef[(size_t)(off + p) * 4 + 3];
            if (thr <= 0) continue;
            const float *b = pa >= 0 ? xyz + (size_t)s->step_sp_env[(size_t)i * MAX_SP_ENV + pa] * 3 : xyz + (size_t)(sc0 - pa - 1) * 3;
            const double r = sqrt((double)dist2(xyz + (size_t)(sc0 + k) * 3, b)) / thr;
            if (r < ratio) ratio = r;
        }
        for (int q = 0; q < s->step_n_sp_env[i]; ++q) is_special[s->step_sp_env[(size_t)i * MAX_SP_ENV + q]] = 0;
    }
    if (min_clash_ratio) *min_clash_ratio = ratio;
    free(is_special); free(clsA); free(clsB);
    return total;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
