/*
 * mcsce_b200 — C ABI of the B200-native MCSCE side-chain growth path.
 *
 * The reference (THGLab/MCSCE) has no FFI for this path; its seam is a set of Python
 * callables.  Each entry point below names the reference code it replaces (paths are
 * relative to the reference's src/mcsce/).  The Python side of the drop-in
 * (mcsce_b200/core, mcsce_b200/libs) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; mcsce_last_error() gives the text;
 *  - pointers named d_* are device pointers (caller-owned, e.g. torch tensor data_ptr()),
 *    pointers named h_* are host pointers; the library allocates only inside the handle;
 *  - all launches go to the cudaStream_t passed as `stream` (void* here, 0 = default);
 *  - a handle is bound to one device and is not thread-safe (one process per GPU);
 *  - atoms are in "growth order": input (backbone) atoms in file order, then one side-chain
 *    block per residue (libs/libstructure.py:498-514).
 */
#ifndef MCSCE_B200_H
#define MCSCE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCSCE_MAX_CLS 64      /* distinct (sigma, epsilon, clash element) atom classes        */
#define MCSCE_MAX_SP_ENV 16   /* environment atoms per residue scored pair-by-pair            */
#define MCSCE_MAX_CHI 6       /* chi columns in a trial row (KCX carries 6, 5 are applied)    */
#define MCSCE_MAX_CHI_ROT 5   /* chi rotations applied (libs/libscbuild.py:103-206)           */
#define MCSCE_MAX_TMPL 32     /* template atoms (M3L: 31)                                     */
#define MCSCE_MAX_SC 28       /* side-chain atoms (M3L: 26)                                   */

typedef struct mcsce_handle_s *mcsce_handle;
typedef struct mcsce_hpmf_s *mcsce_hpmf;

/* Static tables of one sequence.  Replaces what initialize_func_calc builds
 * (core/side_chain_builder.py:42-107) and prepare_energy_function captures per residue
 * (libs/libenergy.py:27-233): per-atom parameters instead of per-pair arrays, plus a short
 * exact-coefficient list for the pairs within three bonds of each new side chain.          */
typedef struct {
    int32_t n_atoms, n_input, n_res, n_cls, n_types, n_sp_total;
    /* per atom, growth order */
    const int32_t *slot_of_atom;   /* [n_atoms] position in the class-sorted environment array   */
    const int32_t *cls_of_atom;    /* [n_atoms]                                                  */
    const double  *charge;         /* [n_atoms] e; zeros disable Coulomb                         */
    const int32_t *resnum;         /* [n_atoms] residue number (bonded rules are keyed on it, libenergy.py:262-270) */
    /* atom classes */
    const int32_t *cls_start;      /* [n_cls+1] slot offsets                                     */
    const int32_t *cls_active;     /* [(n_res+1)*n_cls] atoms of each class present when residue i grows */
    const double  *cls_s2;         /* [n_cls*n_cls] ((sigma_i+sigma_j)*5)^2  A^2  (libenergy.py:478) */
    const double  *cls_e4;         /* [n_cls*n_cls] 4*sqrt(eps_i*eps_j) kJ/mol (libenergy.py:474-481) */
    const double  *cls_thr;        /* [n_cls*n_cls] 0.6*(r_i+r_j) A (libenergy.py:151-152); 0 = no clash test */
    /* residue templates in use */
    const int32_t *tmpl_n_atoms;   /* [n_types]                                                  */
    const float   *tmpl_xyz;       /* [n_types*MAX_TMPL*3] float32 template, CA at origin        */
    const int32_t *tmpl_sc_pos;    /* [n_types*MAX_SC] template position of the k-th side-chain atom */
    const int32_t *tmpl_n_chi;     /* [n_types] rotations defined                                */
    const int32_t *tmpl_axis;      /* [n_types*MAX_CHI_ROT*2] axis (from,to); `to` is the pivot  */
    const uint32_t *tmpl_move;     /* [n_types*MAX_CHI_ROT] bit a set = template atom a moves    */
    const double  *tmpl_ori;       /* [n_types*MAX_CHI_ROT] template dihedral, rad               */
    /* per residue */
    const int32_t *step_kind;      /* [n_res] 0 retained, 1 rigid (GLY/ALA), 2 grown             */
    const int32_t *step_type;      /* [n_res] template index                                     */
    const int32_t *step_n_sc;      /* [n_res]                                                    */
    const int32_t *step_sc_start;  /* [n_res] growth index of the first side-chain atom          */
    const int32_t *step_n_sp_env;  /* [n_res]                                                    */
    const int32_t *step_sp_env;    /* [n_res*MAX_SP_ENV] growth indices of the special environment atoms */
    const int32_t *step_sp_off;    /* [n_res] offset into sp_*                                   */
    const int32_t *step_n_sp;      /* [n_res]                                                    */
    const int32_t *sp_k;           /* [n_sp_total] new-atom index within the side chain          */
    const int32_t *sp_partner;     /* [n_sp_total] >=0: index into the residue's special env list; <0: -(k'+1) */
    const double  *sp_coef;        /* [n_sp_total*4] A, B, Q, clash distance  (exact per-pair values) */
    /* input-atom self energy (core/side_chain_builder.py:149): special pairs of the input atoms */
    int32_t n_bb_pairs;
    const int32_t *bb_i, *bb_j;    /* [n_bb_pairs] growth indices, i<j                           */
    const double  *bb_coef;        /* [n_bb_pairs*4]                                             */
} mcsce_system_desc;

/* Tables that depend on the backbone conformation: trial sets
 * (core/rotamer_library.py:131-177) and the rigid placement of every residue template
 * (libs/libcalc.py:449-482; depends on the backbone only).                                  */
typedef struct {
    int32_t n_trials_total;
    const int32_t *step_n_trials;  /* [n_res]                                                    */
    const int32_t *step_trial_off; /* [n_res]                                                    */
    const int32_t *step_n_chi_lib; /* [n_res] chi columns in the trial rows                      */
    const int32_t *step_n_chi;     /* [n_res] rotations applied                                  */
    const double  *chi_mean;       /* [n_trials_total*MAX_CHI] degrees                           */
    const double  *chi_sigma;      /* [n_trials_total*MAX_CHI] degrees (PTM extra chis), else 0  */
    const double  *prob;           /* [n_trials_total] prior weights                             */
    const double  *place_q1;       /* [n_res*4] quaternion (cos, sin*axis) of the N-CA alignment */
    const double  *place_q2;       /* [n_res*4] quaternion of the N-CA-C plane alignment         */
    const float   *place_ca;       /* [n_res*3] backbone CA                                      */
    const float   *input_xyz;      /* [n_input*3] input atom coordinates (unrounded float32)     */
} mcsce_backbone_desc;

/* Options of one ensemble growth. */
typedef struct {
    int32_t n_conf;                /* conformers grown by this call                              */
    int64_t conf_id0;              /* global id of the first conformer (RNG stream key; rank offset) */
    uint64_t seed;
    double beta;                   /* 1/(kT) mol/kJ (core/side_chain_builder.py:313)             */
    double noise_deg;              /* chi perturbation sigma, 0.5 in the reference (:173)        */
    /* optional replay inputs (parity runs); NULL = on-device Philox */
    const double *d_chis;          /* [n_conf*n_trials_total*MAX_CHI] perturbed chi rows, degrees */
    const double *d_uniform;       /* [n_conf*n_res] uniforms for the selection                  */
    /* optional dumps; NULL = not written */
    double *d_trial_energy;        /* [n_conf*n_trials_total] per-trial energies (inf = clash)   */
    double *d_chis_out;            /* [n_conf*n_trials_total*MAX_CHI] chi rows actually used     */
    int32_t *d_selected;           /* [n_conf*n_res] selected trial per residue (-1 = none)      */
    int32_t env_split;             /* >0 forces the environment split factor; 0 = automatic      */
    int32_t no_dedup;              /* 1: score the chi-independent side-chain atoms (HA, CB, ...) in every trial like the
                                      reference does, instead of once per residue (same result, more work) */
    int32_t no_graph;              /* which driver runs the growth.  0: automatic - launch-bound runs (little work per residue) go to
                                      the persistent cluster kernel (one launch: a thread-block cluster per conformer loops over
                                      the residues on the device), everything else to the per-kernel path; 1: always the per-kernel
                                      path; 2: launch-bound runs replay a cached CUDA graph of the per-kernel path instead;
                                      3: the persistent cluster kernel whatever the size (falls back if the conformer's
                                      environment does not fit one SM's shared memory) */
    int32_t no_screen;             /* 1: score every trial against the whole environment; by default a short clash-only pass over
                                      the atoms near the residue runs first and trials it rejects (+inf either way) are not scored */
    /* optional replay input, used with d_chis (parity runs): the dihedral rotate_sidechain measured before each chi
     * rotation (libs/libscbuild.py:113,167,183,197: float32 NumPy, host dependent), rad; NaN entries and NULL = the
     * template constant.  With the reference's recorded values K1's coordinates are bit-identical to the reference's. */
    const double *d_ori;           /* [n_conf*n_trials_total*MAX_CHI_ROT]                        */
    /* terms = [..., "hpmf"] (libs/libenergy.py:213-220, 806-808): a handle made by mcsce_hpmf_create for this atom list adds
     * the hydrophobic PMF of (conformer so far + trial) to every trial that passed the clash filter, and the PMF of the input
     * atoms alone to the starting energy; NULL = the term is off.  d_ok = 2 marks a conformer given up because one of the
     * term's neighbour-list capacities was exceeded (not expected for physical structures). */
    struct mcsce_hpmf_s *hpmf;
} mcsce_grow_opts;

const char *mcsce_last_error(void);
int mcsce_version(void);

int mcsce_create(mcsce_handle *out, int device);
int mcsce_destroy(mcsce_handle h);

/* replaces initialize_func_calc (core/side_chain_builder.py:42-107) */
int mcsce_set_system(mcsce_handle h, const mcsce_system_desc *h_desc);
/* per-backbone part: rotamer trial sets + template placement (side_chain_builder.py:144-147,171) */
int mcsce_set_backbone(mcsce_handle h, const mcsce_backbone_desc *h_desc);

/* bytes of caller-provided device workspace mcsce_grow needs for n_conf conformers */
size_t mcsce_workspace_bytes(mcsce_handle h, int n_conf);

/* Fused growth of n_conf independent conformers: for every residue, place the trial rotamers,
 * score them against the conformer's environment, pick one by Boltzmann weight, append it.
 * Replaces create_side_chain_structure (core/side_chain_builder.py:109-205) for a batch.
 *   d_coords  [n_conf*n_atoms*3] float32, growth order (out)
 *   d_energy  [n_conf] float64 accumulated energy incl. the input-atom self energy (out)
 *   d_ok      [n_conf] uint8, 0 = every trial of some residue clashed (out)
 * n_conf <= 65535 per call (the Python side chunks larger requests).                          */
int mcsce_grow(mcsce_handle h, const mcsce_grow_opts *opts, void *d_workspace, size_t workspace_bytes,
               float *d_coords, double *d_energy, uint8_t *d_ok, void *stream);

/* Mode reductions of the callers on the device: "exhaustive" = the successful conformer of lowest energy
 * (core/side_chain_builder.py:280-281; ties -> lowest index, like np.argmin), "simple" / return_first_valid = the first
 * successful one (:245-249).  d_energy / d_ok / d_coords as mcsce_grow wrote them; *d_index = -1 when no conformer
 * succeeded; d_picked_coords [n_atoms*3] (optional) receives the picked conformer, so one conformer instead of the
 * ensemble travels to the host.                                                                                     */
int mcsce_pick(mcsce_handle h, int n_conf, const double *d_energy, const uint8_t *d_ok, const float *d_coords,
               int first_valid, int32_t *d_index, float *d_picked_coords, void *stream);

/* Same with HOST result buffers: copies results device->host on `stream` and synchronises.
 * Needs workspace_bytes(n_conf) + n_conf*(n_atoms*12+9) extra bytes of workspace.            */
int mcsce_grow_host(mcsce_handle h, const mcsce_grow_opts *opts, void *d_workspace, size_t workspace_bytes,
                    float *h_coords, double *h_energy, uint8_t *h_ok, void *stream);

/* K1 - chi angles -> placed side-chain coordinates.  Replaces rotate_sidechain
 * (libs/libscbuild.py:56-206) + place_sidechain_template (libs/libcalc.py:419-482) for a batch.
 *   d_chis  [n_rows*MAX_CHI] degrees;  d_xyz [n_rows*n_sc*4] float32 (x,y,z,0) out            */
int mcsce_place_trials(mcsce_handle h, int residue, int n_rows, const double *d_chis, float *d_xyz, void *stream);
/* same with the measured dihedrals replayed: d_ori [n_rows*MAX_CHI_ROT] rad (NaN / NULL = template constant) */
int mcsce_place_trials_replay(mcsce_handle h, int residue, int n_rows, const double *d_chis, const double *d_ori,
                              float *d_xyz, void *stream);

/* K2 - per-trial energy of residue `residue`'s side chain against an environment.  Replaces the
 * closure returned by prepare_energy_function (libs/libenergy.py:786-810) incl. the clash filter.
 *   n_env_sets environments of n_atoms slots each (float4 x,y,z,q; class-sorted slot order),
 *   trials_per_env trial side chains per environment:
 *   d_trial [n_env_sets*trials_per_env*n_sc*4] float32;  d_env [n_env_sets*n_atoms*4] float32
 *   d_energy [n_env_sets*trials_per_env] float64 out, +inf where the clash filter fires       */
int mcsce_score_trials(mcsce_handle h, int residue, int n_env_sets, int trials_per_env, const float *d_trial,
                       const float *d_env, double *d_energy, void *d_workspace, size_t workspace_bytes,
                       int env_split, void *stream);

size_t mcsce_score_workspace_bytes(mcsce_handle h, int n_env_sets, int trials_per_env);

/* growth-order coordinates -> environment arrays for mcsce_score_trials:
 *   d_old_xyz [n_env_sets*n_old*3] float32 (the first n_old atoms in growth order)
 *   d_env     [n_env_sets*n_atoms*4] float32 out (slot order, charge in w; other slots parked far away) */
int mcsce_pack_env(mcsce_handle h, int n_env_sets, int n_old, const float *d_old_xyz, float *d_env, void *stream);

/* K3 - Boltzmann selection.  Replaces side_chain_builder.py:183-193 + choose_random (:33-40).
 *   d_energy [n_sets*n_trials], d_prob [n_trials], d_uniform [n_sets] ->
 *   d_selected [n_sets] (-1 if every trial is +inf), d_emin [n_sets]                           */
int mcsce_select(mcsce_handle h, int n_sets, int n_trials, const double *d_energy, const double *d_prob,
                 double beta, const double *d_uniform, int32_t *d_selected, double *d_e_selected, void *stream);

/* energy of the input atoms alone (core/side_chain_builder.py:149); uses the set backbone */
int mcsce_input_energy(mcsce_handle h, double *d_energy, void *stream);

/* scatter input + rigid atoms into environment arrays: d_env [n_env_sets*n_atoms*4] float32 */
int mcsce_init_env(mcsce_handle h, int n_env_sets, float *d_env, void *stream);

/* counters for the benchmark since the last reset: kernels launched, algorithmic pair evaluations (counted as the
 * reference counts them, libs/libcalc.py:56) and pair evaluations actually executed (fewer: chi-independent atoms
 * are scored once per residue) */
int mcsce_counters(mcsce_handle h, int64_t *launches, int64_t *pairs, int64_t *pairs_computed, int reset);

/* Stand-alone mirrors of the two per-trial reference functions (HOST buffers; no handle needed):
 *   rotate_sidechain(res_type, tors)  (libs/libscbuild.py:56-206): template (n_atoms x 3 float32), rotation
 *   program (axis pairs, moving-atom bit masks, template dihedrals) and n_rows chi rows [n_rows*MAX_CHI] degrees
 *   -> h_out [n_rows*n_atoms*3] float32, template frame;
 *   place_sidechain_template(bb, tmpl) (libs/libcalc.py:419-482): the two placement quaternions (host tables,
 *   mcsce_b200/libs/placement.py) applied to n_atoms float32 points -> float64. */
int mcsce_rotate_template(int device, int n_atoms, const float *h_tmpl_xyz, int n_rot, const int32_t *h_axis,
                          const uint32_t *h_move, const double *h_ori, int n_rows, const double *h_chis, float *h_out);
int mcsce_place_template(int device, int n_atoms, const float *h_xyz, const double *h_q1, const double *h_q2,
                         const float *h_ca, double *h_out);

/* Batched PDB output (host, multi-threaded): replaces the per-conformer Structure.write_PDB of
 * create_side_chain_ensemble (core/side_chain_builder.py:200-204, 321-338; libs/libpdb.py:178-195).
 *   h_coords [n_conf*n_atoms*3] float32 growth order; h_order[i] = growth index of output line i;
 *   line_template: n_atoms lines of line_len bytes (newline included), coordinate fields blank at
 *   [coord_off, coord_off+24); writes <dir>/<n>.pdb for every n with h_ok[n] != 0 (h_ok may be NULL). */
int mcsce_write_pdbs(int n_conf, int n_atoms, const float *h_coords, const uint8_t *h_ok, const int32_t *h_order,
                     const char *line_template, int line_len, int coord_off, const char *dir, int n_threads,
                     int32_t *n_written);

/* Peer-memory result buffer (multi-GPU, one process per GPU): rank 0 allocates the ensemble's result buffer and publishes
 * its 64-byte CUDA IPC handle; the other ranks map it and hand their slice to mcsce_grow as d_coords/d_energy/d_ok, so the
 * result kernels store straight into rank 0's memory over NVLink - no collective, no staging copy.  Replaces the result
 * hand-back of multiprocessing.Pool.imap_unordered (core/side_chain_builder.py:328-347).
 *   alloc: cudaMalloc + cudaIpcGetMemHandle on the owning rank;  open/close: map/unmap on a peer;  free: owner. */
int mcsce_peer_alloc(int device, size_t bytes, void **d_ptr, unsigned char *handle64);
int mcsce_peer_open(int device, const unsigned char *handle64, void **d_ptr);
int mcsce_peer_close(int device, void *d_ptr);
int mcsce_peer_free(int device, void *d_ptr);

/* per-kernel-class device timing for the benchmark's roofline: enable=1 starts recording CUDA-event
 * pairs around every launch (on the launch stream); enable=0 synchronises and returns summed
 * milliseconds / launch counts for [place, score_generic, score_special, select, hpmf] (5 entries each). */
int mcsce_profile(mcsce_handle h, int enable, double *ms_by_class, int64_t *launches_by_class);


/* ------------------------------------------------------------------------------------------------
 * A7 - hydrophobic potential of mean force (libs/libhpmf.py) on the Hasel-Hendrickson-Still approximate
 * solvent-accessible surface (libs/libsasa.py), evaluated for batches of whole structures.
 * The reference cannot reach this term through prepare_energy_function (libs/libenergy.py:110-113 vs :217,
 * :807 vs libhpmf.py:92); the entry points mirror the two closures that do run when called directly.
 * ------------------------------------------------------------------------------------------------ */
#define MCSCE_SASA_MAX_BONDS 4   /* covalent neighbours that carry a SASA type (N-terminal N: CA H1 H2 H3) */

typedef struct {
    int32_t n_atoms;               /* atoms per structure (typed or not)                          */
    int32_t n_typed;               /* atoms with a SASA type (libhpmf.py:22-68: carbon-bound H have none) */
    int32_t n_types;
    const int32_t *typed_atom;     /* [n_typed] index of the typed atom within the structure      */
    const int32_t *type_of;        /* [n_typed] SASA type index                                   */
    const int32_t *bonded;         /* [n_typed*MCSCE_SASA_MAX_BONDS] typed indices of the covalent neighbours, -1 = none
                                      (the bond_connectivities matrix of libsasa.py:76, row by row) */
    const double *type_R;          /* [n_types] radius, A   (libsasa.py:46-70)                     */
    const double *type_P;          /* [n_types] P_i         (libsasa.py:21-45)                     */
    const uint8_t *type_is_carbon; /* [n_types] type name starts with "C" (libhpmf.py:90)          */
    double r_solvent;              /* 1.4 A (libsasa.py:76)                                        */
    double pij_bonded, pij_non_bonded;   /* 0.8875, 0.3516 (libsasa.py:73-74)                      */
    double hpmf_c[3], hpmf_w[3], hpmf_h[3];  /* Gaussian centres, widths, heights (libhpmf.py:17-19) */
    double sa_threshold;           /* carbons with SA <= threshold drop out (libhpmf.py:20, 98)    */
} mcsce_hpmf_desc;

/* replaces libsasa.calc_sasa(...) / libhpmf.init_hpmf_calculator(...): everything that does not depend on coordinates */
int mcsce_hpmf_create(mcsce_hpmf *out, int device, const mcsce_hpmf_desc *h_desc);
int mcsce_hpmf_destroy(mcsce_hpmf h);

/* Atoms with x > 1e14 count as absent (the growth parks unplaced atoms at 1e15): surface area 0, no pair terms -
 * a partially grown structure is evaluated as if it held only the atoms placed so far.
 *
 * replaces the closure of calc_sasa (libsasa.py:89-108) for n_struct structures:
 *   d_coords [n_struct*n_atoms*3] float32  ->  d_sasa [n_struct*n_typed] float64, A^2            */
int mcsce_sasa(mcsce_hpmf h, int n_struct, const float *d_coords, double *d_sasa, void *stream);

/* replaces the closure of init_hpmf_calculator (libhpmf.py:92-109) for n_struct structures:
 *   d_coords [n_struct*n_atoms*3] float32  ->  d_v [n_struct] float64;
 *   d_sasa [n_struct*n_typed] float64: caller-provided scratch, holds the surface areas on return  */
int mcsce_hpmf_energy(mcsce_hpmf h, int n_struct, const float *d_coords, double *d_sasa, double *d_v, void *stream);

/* The term inside the growth (libenergy.py:806-808: coordinate-based terms are added to every trial that passed the
 * clash filter, evaluated on the level's whole structure): V_hpmf of (the atoms before first_new + one trial's new
 * atoms) for every (conformer, trial) with d_mask != 0, without re-evaluating the environment per trial.
 *   d_coords [n_conf*n_atoms*3] float32: the conformers so far; atoms from first_new on are ignored
 *   d_trial  [n_conf*n_trials*n_new*3] float32: the trials' coordinates of atoms [first_new, first_new+n_new)
 *   d_mask   [n_conf*n_trials] uint8 or NULL (= all);  d_v [n_conf*n_trials] float64 out, 0 where masked
 * Synchronises the stream (reads back an overflow flag).  n_conf <= 65535.                               */
int mcsce_hpmf_trials(mcsce_hpmf h, int n_conf, const float *d_coords, int first_new, int n_new, int n_trials,
                      const float *d_trial, const uint8_t *d_mask, double *d_v, void *stream);

/* Growth loops call mcsce_hpmf_trials residue after residue on the same d_coords buffer, each call's environment
 * being the previous one's plus the atoms placed since.  enable=1 lets the handle keep the environment's areas,
 * weights and pair fields between such calls and extend them by the new atoms instead of re-deriving them (the
 * caller promises: same buffer and n_conf, first_new never decreases, atoms below the previous first_new unchanged);
 * every call of this function - with 0 or 1 - forgets the kept state, so a new growth starts with one.  */
int mcsce_hpmf_carry(mcsce_hpmf h, int enable);

/* diagnostics: key 0 = which kernel evaluates the trials inside mcsce_grow (1, default: one block per conformer, a warp per
 * trial; 0: one block per (trial, conformer) - the kernel behind mcsce_hpmf_trials); same numbers to 1e-13 */
int mcsce_hpmf_set_option(mcsce_hpmf h, int key, int value);

/* per-kernel device time of the last mcsce_hpmf_energy call: ms[0] = SASA kernel, ms[1] = carbon-pair kernel
 * (CUDA events on the launch stream; synchronises) */
int mcsce_hpmf_last_ms(mcsce_hpmf h, double *ms);

#ifdef __cplusplus
}
#endif
#endif
