// K0 grow_cluster_kernel - the whole growth of a conformer inside ONE kernel: a thread-block cluster (1-16 CTAs) per
// conformer loops over the residues on the device (A8, core/side_chain_builder.py:109-205; SURVEY §8 A8 / §7 hard part 3).
// Included by mcsce_kernels.cu (same translation unit: it reuses K1's rotation code, K2's packed pair loop and the
// special-pair arithmetic verbatim, so the numerics contract is the one stated there).
//
// Why: a growth is a chain of dependent steps (place -> score -> select -> place ...).  As separate kernels that is 4-5
// launches per residue whose latencies and tails add up when there is little work per residue (few conformers, small
// proteins): config 1 (76 residues, 1 conformer) ran 1.5 ms as a replayed CUDA graph with the SMs 0.4 % busy.  Here
//   * the conformer's environment lives in shared memory for the whole growth, already in the pair-interleaved layout
//     the pair loop reads (x0 x1 y0 y1 | z0 z1 q0 q1, every class segment starting at an even position): no staging, no
//     per-tile barriers - a residue's active environment is a prefix of every class segment, slots not placed yet hold a
//     parked dummy that contributes exactly 0;
//   * the trials of a residue are divided among the CTAs of the cluster; every CTA places and scores its own trials
//     (warps take (64 trial atoms) x (share of the environment) units, no barrier inside the scoring phase), pushes one
//     (energy, clash) per trial into every CTA's shared memory (distributed shared memory), and after ONE cluster barrier
//     per residue every CTA runs the selection redundantly on the same numbers (deterministic, so all CTAs append the same
//     side chain to their copy of the environment);
//   * the loop's tables (step record, class prefix, charges, prior weights, this CTA's special-pair sums and trial atoms) of
//     the NEXT residue are copied into a shared-memory stage with LDGSTS in the shadow of that barrier: a 1-conformer run is
//     bound by dependent-read latency, not arithmetic;
//   * placing the trials (K1) and their special pairs (K2s: own backbone, previous C, next N, the side chain itself) depend
//     on the backbone only, not on what was grown before - so they leave the residue loop: every CTA first places ALL trials
//     of ALL residues it owns and scores their special pairs with every warp busy, then the loop itself is only
//     generic pairs -> barrier -> selection -> append;
//   * random numbers are the same Philox streams as the per-kernel path, keyed by (seed, global conformer id, residue,
//     trial, chi): both paths grow the same conformers up to float32 summation order.
// One launch per call; conformers beyond the resident clusters are taken in a grid-stride loop.

#define GC_MAX_THREADS 512
#define GC_BATCH 2048              // trial atoms scored between two block barriers
#define GC_SE 4096                 // per-atom partial sums held between them (atoms x environment shares)
#define GC_K1_ROWS 5               // K1: trial rows a warp takes at a time (5 rows x 6 chi columns = 30 lanes of scalar work)
#define GC_TRIAL_PF 512            // trial atoms of the next residue staged in shared memory ahead of time (more: read from L2)

// What the residue loop needs of one residue.  It is copied into shared memory one residue ahead with asynchronous copies
// (LDGSTS) issued in the shadow of the cluster barrier, so the loop's dependent table reads are shared-memory reads instead
// of a chain of L2 round trips (~1000 clocks each; measured: half of the scoring phase of a 1-conformer run).
struct __align__(16) GcStage {
    StepDev st;
    int prefix[MAX_CLS + 1];       // row of cls_prefix: atoms of every class present when the residue grows
    int cls[MAX_SC], slot[MAX_SC]; // of its side-chain atoms
    double q[MAX_SC];              // their charges
    int n_tb;                      // trial atoms staged (0: more than GC_TRIAL_PF, read from global memory)
};
static_assert(sizeof(StepDev) % 16 == 0, "StepDev is copied in 16-byte pieces");

__device__ __forceinline__ void gc_ldgsts(void *dst, const void *src, int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src));
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(src));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(src));
}

struct ClusterArgs {
    SysDev s;
    CallParams cp;                 // by value: one launch per call, nothing to replay
    int n_conf, cluster, dedup_ok, max_trials;
    int env_pairs;                 // atom pairs of the resident environment image (even-padded class segments)
    const int *pstart;             // [n_cls+1] first position of every class inside that image (even)
    float4 *trial;                 // [n_conf][trial_stride] placed trials of every residue (residue i at aoff[i])
    long long trial_stride;        // = aoff[n_res]
    const int *aoff;               // [n_res+1] first trial atom of every residue inside a conformer's block
    double2 *sp;                   // [n_conf][n_trials_total] special-pair (energy, clash) of every trial
    const double *chis_in, *ori_in, *uniform;       // replay inputs (null = Philox / template constants)
    double *chis_out, *trial_energy;                // dumps
    int *selected;
    float *coords; double *energy; unsigned char *ok;
    const double *e_input;         // energy of the input atoms alone (set_backbone)
    unsigned long long *pair_counter;
    float unit_overhead;           // fixed cost of a work unit relative to one pass over the whole environment (share heuristic)
    long long *debug;              // optional [16] clock64 ticks per phase (cluster rank 0 of conformer 0; MCSCE_B200_GC_DEBUG)
};

// one (energy, clash) pair into the same shared-memory slot of every CTA of the cluster (distributed shared memory): after
// the cluster barrier every CTA holds all trials' sums locally - no round trip through L2
__device__ __forceinline__ void gc_push(double2 *slot, int cluster, double2 v) {
    if (cluster == 1) { *slot = v; return; }
    const unsigned la = (unsigned)__cvta_generic_to_shared(slot);
    for (int r = 0; r < cluster; ++r) {
        unsigned ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(r));
        asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(ra), "d"(v.x), "d"(v.y) : "memory");
    }
}

// position P of an atom inside the pair-interleaved image -> index of its x in the float view (y +2, z +4, q +6)
__device__ __forceinline__ int gc_flat(int P) { return 8 * (P >> 1) + (P & 1); }

// generic pairs of this thread's RR trial atoms against share j of esplit of the active environment: the active pairs of
// all classes form one sequence (s_pp = its per-class prefix), a share is a contiguous run of it, so a unit walks only the
// classes its run touches
template <int RR>
__device__ __forceinline__ void gc_score_unit(const float4 *__restrict__ s_env, const int *s_pstart,
                                              const int *s_pp, const float *s_cthr, const double *s_cA, const double *s_cB,
                                              int n_cls, int j, int esplit, const float2 (&nx)[SCORE_R],
                                              const float2 (&ny)[SCORE_R], const float2 (&nz)[SCORE_R], const int (&ci)[SCORE_R],
                                              double (&e_lj)[SCORE_R], double (&e_c)[SCORE_R], bool (&clash)[SCORE_R]) {
    const int total = s_pp[n_cls];
    const int g0 = (int)(((unsigned)total * (unsigned)j) / (unsigned)esplit), g1 = (int)(((unsigned)total * (unsigned)(j + 1)) / (unsigned)esplit);
    int c = 0;
    if (esplit > 1) { int c_lo = 0, c_hi = n_cls; while (c_hi - c_lo > 1) { const int m = (c_lo + c_hi) >> 1; if (s_pp[m] <= g0) c_lo = m; else c_hi = m; } c = c_lo; }
    for (; c < n_cls; ++c) {
        const int b = s_pp[c];
        if (b >= g1) break;
        const int lo = max(g0, b), hi = min(g1, s_pp[c + 1]);
        if (lo >= hi) continue;
        const int P0 = s_pstart[c] >> 1;
        const int q0 = P0 + lo - b, q1 = P0 + hi - b;
        float thr2[SCORE_R];
        double A[SCORE_R], B[SCORE_R], E6[SCORE_R], E3[SCORE_R], EC[SCORE_R];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int idx = ci[r] * n_cls + c;
            thr2[r] = s_cthr[idx]; A[r] = s_cA[idx]; B[r] = s_cB[idx];
            E6[r] = 0.0; E3[r] = 0.0; EC[r] = 0.0;
        }
        score_segment<RR>(s_env, q0, q1, nx, ny, nz, thr2, E6, E3, EC, clash);
#pragma unroll
        for (int r = 0; r < RR; ++r) { e_lj[r] += A[r] * E6[r] - B[r] * E3[r]; e_c[r] += EC[r]; }
    }
}

#define GC_TICK(k) do { if (dbg) { const long long t_ = clock64(); a.debug[k] += t_ - t_last; t_last = t_; } } while (0)

__global__ void __launch_bounds__(GC_MAX_THREADS, 1) grow_cluster_kernel(ClusterArgs a) {
    extern __shared__ float4 gc_dyn[];
    __shared__ int s_pstart[MAX_CLS + 1], s_pp[MAX_CLS + 1], s_cstart[MAX_CLS + 1];
    __shared__ GcStage s_stage[2];
    __shared__ float4 s_save[MAX_SP_ENV];
    __shared__ int s_savepos[MAX_SP_ENV];
    __shared__ double s_red[GC_MAX_THREADS / 32];
    __shared__ double s_k1[GC_MAX_THREADS / 32][2][GC_K1_ROWS][MAX_ROT];      // per warp: half-angle sin / cos of the rows in hand
    __shared__ int s_sel, s_nact_next;

    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int S = a.cluster;
    const int lgS = 31 - __clz(S);                  // (cluster sizes are powers of two: trial shares by shifts, no 64-bit divisions)
    const int rank = blockIdx.x & (S - 1);
    const int n_clusters = gridDim.x >> lgS;
    const int n_cls = a.s.n_cls, n_atoms = a.s.n_atoms;
    const int Tmax = a.max_trials;

    // dynamic shared memory
    float4 *s_env = gc_dyn;                                             // [2*env_pairs]
    float *s_flat = reinterpret_cast<float *>(s_env);
    double *s_e = reinterpret_cast<double *>(s_env + 2 * a.env_pairs);  // [GC_SE] per-atom energies of a batch, one per environment share
    const int Tpad = (Tmax + 1) & ~1;                                   // keeps every array below 16-byte aligned
    double *s_E = s_e + GC_SE;                                          // [Tpad] per-trial energies (selection)
    double *s_w = s_E + Tpad;                                           // [Tpad] Boltzmann weights
    double *s_prob2 = s_w + Tpad;                                       // [2][Tpad] staged prior weights
    double2 *s_sp2 = reinterpret_cast<double2 *>(s_prob2 + 2 * Tpad);   // [2][Tpad] staged special-pair sums of this CTA's trials
    double2 *s_part2 = s_sp2 + 2 * Tpad;                                // [2][Tpad+2] (energy, clash) of every trial of the residue, pushed
                                                                        //     by the CTAs that scored them; slot T: chi-independent atoms
    float4 *s_tb2 = reinterpret_cast<float4 *>(s_part2 + 2 * (Tpad + 2)); // [2][GC_TRIAL_PF] staged trial atoms of this CTA's trials
    double *s_cA = reinterpret_cast<double *>(s_tb2 + 2 * GC_TRIAL_PF); // [n_cls^2] class-pair constants (A/d^12 - B/d^6)
    double *s_cB = s_cA + n_cls * n_cls;
    float *s_cthr = reinterpret_cast<float *>(s_cB + n_cls * n_cls);    // [n_cls^2] exact float32 clash thresholds (squared)
    unsigned char *s_cl = reinterpret_cast<unsigned char *>(s_cthr + ((n_cls * n_cls + 3) & ~3));   // [GC_SE]
    unsigned char *s_kind = s_cl + GC_SE;                               // [n_res] 2 = grown here, 1 = rigid side chain, 0 = nothing to do
    double *s_red2 = s_e + GC_SE - GC_BATCH;                            // per-atom sums over the environment shares: the tail of s_e (a batch
    unsigned char *s_cl2 = s_cl + GC_SE - GC_BATCH;                     //   with more than one share fills at most GC_SE - GC_BATCH entries)

    for (int c = tid; c <= n_cls; c += nthr) { s_pstart[c] = a.pstart[c]; s_cstart[c] = a.s.cls_start[c]; }
    for (int i = tid; i < n_cls * n_cls; i += nthr) { s_cA[i] = a.s.cls_A[i]; s_cB[i] = a.s.cls_B[i]; s_cthr[i] = a.s.cls_thr2f[i]; }
    for (int i = tid; i < a.s.n_res; i += nthr) {
        const int k = a.s.steps[i].kind;
        s_kind[i] = (unsigned char)((k == 2 && a.s.steps[i].n_trials <= 0) ? 0 : k);
    }
    const bool dbg = a.debug != nullptr && blockIdx.x == 0 && tid == 0;
    long long t_last = dbg ? clock64() : 0;

    for (int set = blockIdx.x >> lgS; set < a.n_conf; set += n_clusters) {
        __syncthreads();
        // ---- the conformer's environment: input atoms now, side chains as they are grown (parked dummies until then) ----
        for (int p = tid; p < 2 * a.env_pairs; p += nthr)
            s_env[p] = (p & 1) ? make_float4(1.0e15f, 1.0e15f, 0.f, 0.f) : make_float4(1.0e15f, 1.0e15f, 1.0e15f, 1.0e15f);
        __syncthreads();
        for (int atom = tid; atom < a.s.n_input; atom += nthr) {
            const int c = a.s.cls_of_atom[atom];
            const int f = gc_flat(s_pstart[c] + a.s.slot_of_atom[atom] - a.s.cls_start[c]);
            s_flat[f] = a.s.input_xyz[atom * 3]; s_flat[f + 2] = a.s.input_xyz[atom * 3 + 1];
            s_flat[f + 4] = a.s.input_xyz[atom * 3 + 2]; s_flat[f + 6] = (float)a.s.charge[atom];
        }
        double e_acc = a.e_input[0];                       // starts from the input atoms' energy (side_chain_builder.py:149)
        unsigned long long n_pairs = 0, n_pairs_done = 0;
        bool alive = true;
        int count = 0;
        const long long chi_set_stride = (long long)a.s.n_trials_total * MAX_CHI;
        const long long cid = a.cp.conf_id0 + set;
        float4 *tb_all = a.trial + (long long)set * a.trial_stride;
        double2 *sp_all = a.sp + (long long)set * a.s.n_trials_total;
        __syncthreads();
        GC_TICK(0);

        // ---- K1 for every residue: chi rows of this CTA's trials -> placed coordinates (arithmetic of place_trials_kernel).
        //      A warp takes GC_K1_ROWS rows at a time: lanes (row, chi column) do the scalar part (chi value, half-angle
        //      sin/cos), then lane = template atom rotates and places the rows one after another.  No block barrier. ----
        for (int step = 0; step < a.s.n_res; ++step) {
            const StepDev &st = a.s.steps[step];
            if (st.kind != 2 || st.n_trials <= 0) continue;
            const int T = st.n_trials, n_sc = st.n_sc;
            const int t_lo = ((T * rank) >> lgS), t_hi = ((T * (rank + 1)) >> lgS);
            const TmplDev &tm = a.s.tmpl[st.type];
            float4 *tb = tb_all + a.aoff[step];
            const bool has = lane < tm.n_atoms;
            float x0 = 0.f, y0 = 0.f, z0 = 0.f;
            if (has) { x0 = tm.xyz[lane][0]; y0 = tm.xyz[lane][1]; z0 = tm.xyz[lane][2]; }
            const int sc_rank = has ? tm.sc_rank[lane] : -1;
            const int w_rot = (warp + step) % nwarps;       // rotate the first warp so that short residues spread over the warps
            for (int b0 = t_lo + GC_K1_ROWS * w_rot; b0 < t_hi; b0 += GC_K1_ROWS * nwarps) {
                const int nb = min(GC_K1_ROWS, t_hi - b0);
                {
                    const int tl = lane / MAX_CHI, c = lane - tl * MAX_CHI, t = b0 + tl;
                    if (tl < nb) {
                        const long long row = (long long)st.trial_off + t;
                        const long long at = (long long)set * chi_set_stride + row * MAX_CHI + c;
                        double v = 0.0;
                        if (a.chis_in) {
                            v = a.chis_in[at];
                        } else if (c < st.n_chi_lib) {
                            // wrap(mean + sigma*z1) + noise*z2   (rotamer_library.py:155,165 ; side_chain_builder.py:173)
                            const unsigned long long seed = a.cp.seed;
                            uint4 r = philox4x32(make_uint4((unsigned)cid, (unsigned)(cid >> 32), (unsigned)step,
                                                            ((unsigned)t << 3) | (unsigned)c),
                                                 make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
                            double u1 = u01_open(r.x, r.y), u2 = u01_halfopen(r.z, r.w);
                            double rad = sqrt(-2.0 * log(u1)), sn, cs;
                            sincospi(2.0 * u2, &sn, &cs);
                            v = a.s.chi_mean[row * MAX_CHI + c];
                            const double sg = a.s.chi_sigma[row * MAX_CHI + c];
                            if (sg != 0.0) {
                                v = v + sg * (rad * cs);
                                if (v > 180.0) v -= 360.0;
                                if (v < -180.0) v += 360.0;
                            }
                            v = v + a.cp.noise_deg * (rad * sn);
                        }
                        if (a.chis_out) a.chis_out[at] = v;
                        if (c < st.n_chi) {
                            double ori = tm.ori[c];
                            if (a.ori_in) {                // replay of the reference's measured dihedrals (DESIGN "K1 fidelity")
                                const double m = a.ori_in[((long long)set * a.s.n_trials_total + row) * MAX_ROT + c];
                                if (m == m) ori = m;
                            }
                            chi_half_sincos(ori, v, s_k1[warp][0][tl][c], s_k1[warp][1][tl][c]);
                        }
                    }
                }
                __syncwarp();
                for (int tl = 0; tl < nb; ++tl) {
                    float x = x0, y = y0, z = z0;
                    apply_chi_rotations(tm, st.n_chi, s_k1[warp][0][tl], s_k1[warp][1][tl], lane, has, x, y, z);
                    if (sc_rank >= 0) {                // rigid placement: two rotations (first stored to float32) + CA (libcalc.py:465,480-482)
                        double ox, oy, oz;
                        quat_rotate(st.q1[0], st.q1[1], st.q1[2], st.q1[3], (double)x, (double)y, (double)z, ox, oy, oz);
                        const float rx = __double2float_rn(ox), ry = __double2float_rn(oy), rz = __double2float_rn(oz);
                        quat_rotate(st.q2[0], st.q2[1], st.q2[2], st.q2[3], (double)rx, (double)ry, (double)rz, ox, oy, oz);
                        float4 o;
                        o.x = __double2float_rn(__dadd_rn(ox, (double)st.ca[0]));
                        o.y = __double2float_rn(__dadd_rn(oy, (double)st.ca[1]));
                        o.z = __double2float_rn(__dadd_rn(oz, (double)st.ca[2]));
                        o.w = 0.f;
                        tb[(long long)(b0 + tl) * n_sc + sc_rank] = o;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        GC_TICK(1);

        // ---- K2s for every residue: pairs within three bonds + intra-side-chain pairs of this CTA's trials
        //      (score_special_kernel); the environment partners are input atoms, present from the start ----
        for (int step = 0; step < a.s.n_res; ++step) {
            const StepDev &st = a.s.steps[step];
            if (st.kind != 2 || st.n_trials <= 0) continue;
            const int T = st.n_trials, n_sc = st.n_sc;
            const int t_lo = ((T * rank) >> lgS), t_hi = ((T * (rank + 1)) >> lgS), nt = t_hi - t_lo;
            const float4 *tb = tb_all + a.aoff[step];
            const int w_rot = (warp + step) % nwarps;
            for (int g = w_rot; g * SP_TPW < nt; g += nwarps) {
                const int tb0 = g * SP_TPW;                                     // local trial index
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
                    if (pa >= 0) {
                        const int slot = st.sp_env_slot[pa];
                        int c_lo = 0, c_hi = n_cls;                            // largest c with cls_start[c] <= slot
                        while (c_hi - c_lo > 1) { const int m = (c_lo + c_hi) >> 1; if (a.s.cls_start[m] <= slot) c_lo = m; else c_hi = m; }
                        const int f = gc_flat(s_pstart[c_lo] + slot - a.s.cls_start[c_lo]);
                        bo_env = make_float4(s_flat[f], s_flat[f + 2], s_flat[f + 4], 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < SP_TPW; ++i) {
                        if (tb0 + i < nt) {
                            const float4 *trial = tb + (long long)(t_lo + tb0 + i) * n_sc;
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
                    if (lane == 0 && tb0 + i < nt) sp_all[st.trial_off + t_lo + tb0 + i] = make_double2(v, cl ? 1.0 : 0.0);
                }
            }
        }
        __syncthreads();
        GC_TICK(2);

        // ---- the residue loop: generic pairs -> one cluster barrier -> selection -> append ----
        // stage_issue: asynchronous copies of everything the loop reads of residue nx into stage buffer b (all threads)
        auto stage_issue = [&](int b, int nx) {
            const StepDev *gs = a.s.steps + nx;
            const int n_sc_n = gs->n_sc, sc0 = gs->sc_start, Tn = gs->n_trials, toff = gs->trial_off;
            const int tn_lo = ((Tn * rank) >> lgS), ntn = ((Tn * (rank + 1)) >> lgS) - tn_lo;
            GcStage &sg = s_stage[b];
            const int n_tb = (ntn * n_sc_n <= GC_TRIAL_PF) ? ntn * n_sc_n : 0;
            if (tid == 0) sg.n_tb = n_tb;
            const int o1 = (int)(sizeof(StepDev) / 16), o2 = o1 + n_cls + 1, o3 = o2 + n_sc_n, o4 = o3 + n_sc_n, o5 = o4 + n_sc_n,
                      o6 = o5 + Tn, o7 = o6 + ntn, o8 = o7 + n_tb;
            const float4 *tb_n = tb_all + a.aoff[nx] + (long long)tn_lo * n_sc_n;
            for (int it = tid; it < o8; it += nthr) {
                if (it < o1) gc_ldgsts(reinterpret_cast<char *>(&sg.st) + 16 * it, reinterpret_cast<const char *>(gs) + 16 * it, 16);
                else if (it < o2) gc_ldgsts(&sg.prefix[it - o1], a.s.cls_prefix + nx * (n_cls + 1) + (it - o1), 4);
                else if (it < o3) gc_ldgsts(&sg.cls[it - o2], a.s.cls_of_atom + sc0 + (it - o2), 4);
                else if (it < o4) gc_ldgsts(&sg.slot[it - o3], a.s.slot_of_atom + sc0 + (it - o3), 4);
                else if (it < o5) gc_ldgsts(&sg.q[it - o4], a.s.charge + sc0 + (it - o4), 8);
                else if (it < o6) gc_ldgsts(&s_prob2[b * Tpad + (it - o5)], a.s.prob + toff + (it - o5), 8);
                else if (it < o7) gc_ldgsts(&s_sp2[b * Tpad + (it - o6)], sp_all + toff + tn_lo + (it - o6), 16);
                else gc_ldgsts(&s_tb2[b * GC_TRIAL_PF + (it - o7)], tb_n + (it - o7), 16);
            }
            asm volatile("cp.async.commit_group;");
        };
        // prep: active atoms per class of a staged residue (warp wa) and its special environment atoms - scored pair by pair
        // by K2s - saved and parked for the generic pass (lanes of warp ws)
        auto prep = [&](const GcStage &g, int wa, int ws) {
            if (warp == wa) {                              // (n_cls <= MAX_CLS = 64: two classes per lane)
                int carry = 0;
                for (int c0 = 0; c0 < n_cls; c0 += 32) {
                    const int c = c0 + lane;
                    int act = 0;
                    if (c < n_cls) act = g.prefix[c + 1] - g.prefix[c];
                    const int np = (act + 1) >> 1;
                    int x = np;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
                    if (c < n_cls) s_pp[c] = carry + x - np;
                    carry += __shfl_sync(0xffffffffu, x, 31);
                }
                if (lane == 0) s_pp[n_cls] = carry;
                int na = 0;
                for (int c = lane; c < n_cls; c += 32) na += g.prefix[c + 1] - g.prefix[c];
                na = __reduce_add_sync(0xffffffffu, na);
                if (lane == 0) s_nact_next = na;
            }
            if (warp == ws && lane < g.st.n_sp_env) {
                const int slot = g.st.sp_env_slot[lane];
                int c_lo = 0, c_hi = n_cls;                                    // largest c with cls_start[c] <= slot
                while (c_hi - c_lo > 1) { const int m = (c_lo + c_hi) >> 1; if (s_cstart[m] <= slot) c_lo = m; else c_hi = m; }
                const int f = gc_flat(s_pstart[c_lo] + slot - s_cstart[c_lo]);
                s_save[lane] = make_float4(s_flat[f], s_flat[f + 2], s_flat[f + 4], s_flat[f + 6]);
                s_savepos[lane] = f;
                s_flat[f] = 1.0e15f; s_flat[f + 2] = 1.0e15f; s_flat[f + 4] = 1.0e15f; s_flat[f + 6] = 0.f;
            }
        };
        auto next_grown = [&](int from) { int nx = from; while (nx < a.s.n_res && s_kind[nx] != 2) ++nx; return nx; };
        {
            const int first = next_grown(0);
            if (first < a.s.n_res) stage_issue(0, first);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (first < a.s.n_res) prep(s_stage[0], nwarps - 1, 0);
            __syncthreads();
        }
        GC_TICK(3);
        for (int step = 0; step < a.s.n_res; ++step) {
            const int kind = s_kind[step];
            if (kind == 1) {                               // GLY/ALA: conformer independent, placed by set_backbone (:164-168)
                const StepDev &sr = a.s.steps[step];
                if (tid < sr.n_sc) {
                    const int atom = sr.sc_start + tid, c = a.s.cls_of_atom[atom];
                    const int f = gc_flat(s_pstart[c] + a.s.slot_of_atom[atom] - s_cstart[c]);
                    s_flat[f] = a.s.rigid_xyz[atom * 3]; s_flat[f + 2] = a.s.rigid_xyz[atom * 3 + 1];
                    s_flat[f + 4] = a.s.rigid_xyz[atom * 3 + 2]; s_flat[f + 6] = (float)a.s.charge[atom];
                }
                __syncthreads();
                continue;
            }
            if (kind != 2) continue;
            const int par = count & 1;
            const GcStage &sg = s_stage[par];
            const StepDev &st = sg.st;                     // (shared memory)
            const int n_sc = st.n_sc;
            const int T = st.n_trials;
            const int t_lo = ((T * rank) >> lgS), t_hi = ((T * (rank + 1)) >> lgS), nt = t_hi - t_lo;
            const float4 *tb = tb_all + a.aoff[step];
            // this CTA's trials, local index: staged in shared memory when they fit, else where K1 left them
            const float4 *tloc = sg.n_tb ? s_tb2 + par * GC_TRIAL_PF : tb + (long long)t_lo * n_sc;
            const double2 *sp = s_sp2 + par * Tpad;        // local index as well
            const double *prob = s_prob2 + par * Tpad;
            double2 *part = s_part2 + par * (Tpad + 2);
            const int nx_step = next_grown(step + 1);
            const int n_act = s_nact_next;                 // active environment atoms of this residue (prep)

            // ---- K2: generic pairs of this CTA's trial atoms against the resident environment ----
            const int dedup = (a.dedup_ok && st.n_fix > 0 && st.n_var > 0) ? 1 : 0;
            const int n_per = dedup ? st.n_var : n_sc;
            const int n_fix_here = (dedup && rank == S - 1) ? st.n_fix : 0;     // chi-independent atoms: once per residue
            const int tpb = max(1, (GC_BATCH - MAX_SC) / n_per);
            for (int bt = 0; bt < nt; bt += tpb) {
                const int nbt = min(tpb, nt - bt);
                const bool lastb = bt + tpb >= nt;
                const int n_varatoms = nbt * n_per;
                const int n_fixatoms = lastb ? n_fix_here : 0;
                const int n_local = n_varatoms + n_fixatoms;
                const int chunks = (n_local + 63) >> 6;
                // work units = (64 trial atoms) x (share of every class segment): the number of shares that fills the warps
                // best - time ~ rounds of units x (1/shares + a fixed cost per unit)
                int esplit = 1;
                {
                    float best = 1.0e30f;
                    const int e_max = min(16, (GC_SE - GC_BATCH) / max(1, n_local));
                    const int wshift = 31 - __clz(nwarps);               // (8 or 16 warps)
                    if (chunks == 1) esplit = min(e_max, nwarps);      // one chunk: a share per warp
                    else
                        for (int e = 1, units_e = chunks; e <= e_max; ++e, units_e += chunks) {
                            const float cost = (float)((units_e + nwarps - 1) >> wshift) * (__frcp_rn((float)e) + a.unit_overhead);
                            if (cost < best) { best = cost; esplit = e; }
                            if (units_e >= 4 * nwarps) break;            // beyond four rounds of units more shares only add overhead
                        }
                }
                const int units = chunks * esplit;
                for (int u = warp; u < units; u += nwarps) {
                    const int ch = u / esplit, j = u - ch * esplit;
                    float2 nx[SCORE_R], ny[SCORE_R], nz[SCORE_R];
                    int ci[SCORE_R], katom[SCORE_R];
                    double e_lj[SCORE_R], e_c[SCORE_R];
                    bool clash[SCORE_R];
#pragma unroll
                    for (int r = 0; r < SCORE_R; ++r) {
                        const int la = ch * 64 + r * 32 + lane;
                        int k = 0; float4 p = tloc[0];
                        if (la < n_varatoms) {
                            const int tl = la / n_per, kk = la - tl * n_per;
                            k = dedup ? st.var_k[kk] : kk;
                            p = tloc[(bt + tl) * n_sc + k];
                        } else if (la < n_local) {
                            k = st.fix_k[la - n_varatoms];
                            p = tloc[k];                                         // identical in every trial: this CTA's first
                        }
                        katom[r] = k;
                        nx[r] = make_float2(-p.x, -p.x); ny[r] = make_float2(-p.y, -p.y); nz[r] = make_float2(-p.z, -p.z);
                        ci[r] = sg.cls[k];
                        e_lj[r] = 0.0; e_c[r] = 0.0; clash[r] = false;
                    }
                    if (ch * 64 + 32 < n_local) gc_score_unit<2>(s_env, s_pstart, s_pp, s_cthr, s_cA, s_cB, n_cls, j, esplit, nx, ny, nz, ci, e_lj, e_c, clash);
                    else gc_score_unit<1>(s_env, s_pstart, s_pp, s_cthr, s_cA, s_cB, n_cls, j, esplit, nx, ny, nz, ci, e_lj, e_c, clash);
#pragma unroll
                    for (int r = 0; r < SCORE_R; ++r) {
                        const int la = ch * 64 + r * 32 + lane;
                        if (la < n_local) {
                            const double qi = sg.q[katom[r]];
                            s_e[j * n_local + la] = e_lj[r] + COULOMB_PREF * qi * e_c[r];
                            s_cl[j * n_local + la] = clash[r] ? 1 : 0;
                        }
                    }
                }
                __syncthreads();
                GC_TICK(4);
                // per-atom totals -> per-trial (energy, clash), fixed order; the special pairs join here
                if (esplit > 1) {                          // first the environment shares of every atom: 16 lanes per atom, fixed tree
                    for (int la0 = (tid >> 4); la0 < ((n_local + 1) & ~1); la0 += (nthr >> 4)) {
                        const int j = tid & 15;
                        double e = 0.0; int cl = 0;
                        if (la0 < n_local && j < esplit) { e = s_e[j * n_local + la0]; cl = s_cl[j * n_local + la0]; }
#pragma unroll
                        for (int off = 8; off > 0; off >>= 1) {
                            e += __shfl_xor_sync(0xffffffffu, e, off);
                            cl |= __shfl_xor_sync(0xffffffffu, cl, off);
                        }
                        if (j == 0 && la0 < n_local) { s_red2[la0] = e; s_cl2[la0] = (unsigned char)cl; }
                    }
                    __syncthreads();
                }
                const double *s_ea = esplit > 1 ? s_red2 : s_e;
                const unsigned char *s_ca = esplit > 1 ? s_cl2 : s_cl;
                for (int tl = tid; tl < nbt; tl += nthr) {
                    double e = 0.0; int cl = 0;
                    for (int k = 0; k < n_per; ++k) { e += s_ea[tl * n_per + k]; cl |= s_ca[tl * n_per + k]; }
                    const double2 s2 = sp[bt + tl];
                    e = s2.x + e; cl |= (s2.y != 0.0) ? 1 : 0;
                    gc_push(&part[t_lo + bt + tl], S, make_double2(e, cl ? 1.0 : 0.0));
                }
                if (n_fixatoms > 0 && tid == nthr - 1) {
                    double e = 0.0; int cl = 0;
                    for (int k = 0; k < n_fixatoms; ++k) { e += s_ea[n_varatoms + k]; cl |= s_ca[n_varatoms + k]; }
                    gc_push(&part[T], S, make_double2(e, cl ? 1.0 : 0.0));
                }
                __syncthreads();
            }
            if (tid < st.n_sp_env) {                       // the parked special atoms come back
                const int f = s_savepos[tid];
                const float4 v = s_save[tid];
                s_flat[f] = v.x; s_flat[f + 2] = v.y; s_flat[f + 4] = v.z; s_flat[f + 6] = v.w;
            }

            GC_TICK(5);
            // ---- one barrier per residue: every CTA of the cluster has written its trials' sums.  In its shadow the tables and
            //      trial atoms of the next grown residue start travelling to the other stage buffer ----
            if (S > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            if (nx_step < a.s.n_res) stage_issue(par ^ 1, nx_step);
            if (S > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            else __syncthreads();
            GC_TICK(6);

            // ---- K3: clash mask + Boltzmann selection, by every CTA on the same numbers (side_chain_builder.py:183-197) ----
            double e_fix = 0.0; bool cl_fix = false;
            if (dedup) { const double2 p = part[T]; e_fix = p.x; cl_fix = p.y != 0.0; }
            double mn = CUDART_INF;
            for (int t = tid; t < T; t += nthr) {
                const double2 p = part[t];
                double e = p.x + e_fix;
                if (p.y != 0.0 || cl_fix) e = CUDART_INF;                       // energies[clash] = inf (libenergy.py:801)
                s_E[t] = e;
                if (a.trial_energy && rank == 0) a.trial_energy[(long long)set * a.s.n_trials_total + st.trial_off + t] = e;
                mn = fmin(mn, e);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
            if (lane == 0) s_red[warp] = mn;
            __syncthreads();
            mn = s_red[0];
            for (int w = 1; w < nwarps; ++w) mn = fmin(mn, s_red[w]);
            if (tid == 0) {
                const long long n_active = n_act;
                n_pairs += (unsigned long long)T * ((long long)n_sc * (n_sc - 1) / 2 + (long long)n_sc * n_active);
                n_pairs_done += (unsigned long long)((dedup ? (long long)st.n_fix * n_active : 0) + (long long)T * st.n_sp +
                                                     (long long)T * n_per * n_active);
            }
            if (isinf(mn) || isnan(mn)) {                  // every trial clashes: the growth fails (side_chain_builder.py:187)
                alive = false;
                if (a.selected && rank == 0)
                    for (int s2 = step + tid; s2 < a.s.n_res; s2 += nthr)
                        if (s_kind[s2] == 2) a.selected[(long long)set * a.s.n_res + s2] = -1;
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                break;
            }
            GC_TICK(7);
            for (int t = tid; t < T; t += nthr)
                s_w[t] = prob[t] * exp(-a.cp.beta * (s_E[t] - mn));    // side_chain_builder.py:191-192
            asm volatile("cp.async.wait_group 0;" ::: "memory");          // the next residue's stage buffer has landed
            __syncthreads();
            // while thread 0 walks the running sums, two other warps prepare the next residue's scoring pass (its special
            // atoms are input atoms: parking them now does not touch what the selection appends)
            if (nx_step < a.s.n_res) prep(s_stage[par ^ 1], nwarps - 1, nwarps - 2);
            if (tid == 0) {
                double u;
                if (a.uniform) u = a.uniform[(long long)set * a.s.n_res + step];
                else {
                    const unsigned long long seed = a.cp.seed;
                    uint4 r = philox4x32(make_uint4((unsigned)cid, (unsigned)(cid >> 32), (unsigned)step, 0xFFFFFFFFu),
                                         make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
                    u = u01_halfopen(r.x, r.y);
                }
                // sequential float64 running sums, first index whose sum exceeds u*total (side_chain_builder.py:35-40);
                // the loads run ahead of the dependent additions
                double run = 0.0;
                int t = 0;
                for (; t + 8 <= T; t += 8) {
                    double w8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) w8[q] = s_w[t + q];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { run += w8[q]; s_w[t + q] = run; }
                }
                for (; t < T; ++t) { run += s_w[t]; s_w[t] = run; }
                const double pointer = u * run;
                int sel = T - 1;
                for (t = 0; t + 8 <= T; t += 8) {
                    double w8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) w8[q] = s_w[t + q];
                    int hit = -1;
#pragma unroll
                    for (int q = 7; q >= 0; --q) if (w8[q] > pointer) hit = t + q;
                    if (hit >= 0) { sel = hit; t = T; break; }
                }
                for (; t < T; ++t) if (s_w[t] > pointer) { sel = t; break; }
                s_sel = sel;
                e_acc += (s_E[sel] - mn) + mn;                                  // side_chain_builder.py:197
                if (a.selected && rank == 0) a.selected[(long long)set * a.s.n_res + step] = sel;
            }
            __syncthreads();
            GC_TICK(8);
            if (tid < n_sc) {                              // append the chosen side chain to this CTA's environment
                const int sel = s_sel;
                const float4 p = __ldcg(&tb[(long long)sel * n_sc + tid]);
                const int c = sg.cls[tid];
                const int f = gc_flat(s_pstart[c] + sg.slot[tid] - s_cstart[c]);
                s_flat[f] = p.x; s_flat[f + 2] = p.y; s_flat[f + 4] = p.z; s_flat[f + 6] = (float)sg.q[tid];
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");          // the next residue's stage buffer has landed
            __syncthreads();
            GC_TICK(9);
            ++count;
        }

        // ---- results: growth-order coordinates, accumulated energy, success flag ----
        __syncthreads();
        if (rank == 0) {
            for (int atom = tid; atom < n_atoms; atom += nthr) {
                const int c = a.s.cls_of_atom[atom];
                const int f = gc_flat(s_pstart[c] + a.s.slot_of_atom[atom] - a.s.cls_start[c]);
                float *o = a.coords + ((long long)set * n_atoms + atom) * 3;
                o[0] = s_flat[f]; o[1] = s_flat[f + 2]; o[2] = s_flat[f + 4];
            }
            if (tid == 0) {
                a.ok[set] = alive ? 1 : 0;
                a.energy[set] = alive ? e_acc : CUDART_NAN;
                if (a.pair_counter) { atomicAdd(a.pair_counter, n_pairs); atomicAdd(a.pair_counter + 1, n_pairs_done); }
            }
        }
        // the CTAs of a cluster write into each other's shared memory: none may start the next conformer (or leave) before all
        // have read this one's last sums
        if (S > 1) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
}
