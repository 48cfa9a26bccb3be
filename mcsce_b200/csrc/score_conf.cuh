// K2c score_conf_kernel - the scoring pass (generic pairs of the trials the screen left) with ONE CTA per conformer (and
// environment share): the active environment is staged into shared memory once, in the pair-interleaved layout with every
// class segment even-padded, and the warps then take (64 trial atoms) x (share of the staged pairs) work units with no
// barrier inside the pass - the structure of the persistent kernel's scoring phase (grow_cluster.cuh), as a per-residue
// kernel.  Against score_generic_kernel<false> (a CTA per 256-atom trial tile, each staging the whole environment through
// 1024-atom tiles with two barriers per tile): one staging per conformer instead of one per tile, no partially filled
// tiles idling at barriers, units sized to fill the warps.  Same pair arithmetic (score_segment), same partial-sum layout,
// so the special-pair kernel and the selection are unchanged.  The neighbour screen runs INSIDE this kernel (list_cap_pairs
// > 0): the residue's short list of nearby atoms is staged first, the copies of the whole environment are issued behind it,
// and while those are in flight every trial is tested against the list (same clash_segment, same decision rule, the
// special-pair kernel's flags joined) - one launch, one prologue and one flag round trip less per residue than the
// stand-alone screen pass (score_generic_kernel<1>).  Included by mcsce_kernels.cu.

#define SC_THREADS 256
#define SC_BATCH 1536              // trial atoms scored between two block barriers
#define SC_SE 3072                 // per-atom partial sums held between them (atoms x environment shares)

__global__ void __launch_bounds__(SC_THREADS, 2) score_conf_kernel(ScoreArgs a, int cap_pairs, int fixed_shares, int list_cap_pairs) {
    extern __shared__ float4 sc_dyn[];
    __shared__ int s_pre[MAX_CLS + 1];              // active atoms before every class (row of cls_prefix)
    __shared__ int s_ps[MAX_CLS + 1];               // even-padded start of every class inside the staged share
    __shared__ int s_pp[MAX_CLS + 1];               // the same in pairs (what gc_score_unit walks)
    __shared__ int s_first[MAX_CLS];                // active index of the class's first staged atom
    __shared__ unsigned s_mask[SCREEN_WORDS];
    __shared__ int s_cnt[SCREEN_WORDS + 1];
    __shared__ int s_lpre[MAX_CLS + 1];             // fused screen: atoms of the residue's near list before every class
    __shared__ int s_lps[MAX_CLS + 1];              // even-padded start of every class inside the staged list

    const int set = blockIdx.x, z = blockIdx.y;
    if (a.alive && !a.alive[set]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = SC_THREADS / 32;
    const StepDev &st = a.s.steps[a.step];
    const int n_cls = a.s.n_cls, n_sc = st.n_sc;
    const bool compact = a.screen != nullptr;
    const bool fused = list_cap_pairs > 0;          // the neighbour screen runs here (needs a.screen, a.sp_partial)
    const int dedup = a.dedup;
    const int n_per = dedup ? st.n_var : n_sc;
    const int n_keep_all = (compact && !fused) ? a.n_keep[set] : a.n_trials;
    if (n_keep_all <= 0 && !dedup) return;

    float4 *s_env = sc_dyn;
    float *s_flat = reinterpret_cast<float *>(s_env);
    float4 *s_list = s_env + 2 * cap_pairs;                                       // [2*list_cap_pairs]
    double *s_e = reinterpret_cast<double *>(s_list + 2 * list_cap_pairs);
    unsigned char *s_cl = reinterpret_cast<unsigned char *>(s_e + SC_SE);
    unsigned char *s_flag = s_cl + SC_SE;                                         // [SCREEN_WORDS*32] fused screen: 1 = the trial clashes

    for (int c = tid; c <= n_cls; c += SC_THREADS) {
        s_pre[c] = a.s.cls_prefix[a.step * (n_cls + 1) + c];
        if (fused) s_lpre[c] = a.s.scr_prefix[a.step * (n_cls + 1) + c];
    }
    __syncthreads();
    const int n_active = s_pre[n_cls];
    const int per = (n_active + a.n_split - 1) / a.n_split;
    const int lo = z * per, hi = min(n_active, lo + per);
    if (warp == 0) {                                 // layout of the staged share (n_cls <= 64)
        int carry = 0;
        for (int c0 = 0; c0 < n_cls; c0 += 32) {
            const int c = c0 + lane;
            int len = 0, first = 0;
            if (c < n_cls) { first = max(s_pre[c], lo); len = max(0, min(s_pre[c + 1], hi) - first); }
            const int even = (len + 1) & ~1;
            int x = even;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
            if (c < n_cls) { s_ps[c] = carry + x - even; s_pp[c] = (carry + x - even) >> 1; s_first[c] = first; }
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) { s_ps[n_cls] = carry; s_pp[n_cls] = carry >> 1; }
    }
    if (fused && warp == 1) {                        // layout of the staged near list
        int carry = 0;
        for (int c0 = 0; c0 < n_cls; c0 += 32) {
            const int c = c0 + lane;
            const int len = c < n_cls ? s_lpre[c + 1] - s_lpre[c] : 0;
            const int even = (len + 1) & ~1;
            int x = even;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
            if (c < n_cls) s_lps[c] = carry + x - even;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) s_lps[n_cls] = carry;
    }
    // surviving trials -> compacted index (as in score_generic_kernel); with the fused screen that follows the screen phase
    if (compact && !fused) {
        const unsigned char *flag = a.screen + (long long)set * a.n_trials;
        const int n_words = (a.n_trials + 31) >> 5;
        for (int wb = 0; wb < n_words; wb += nwarps) {
            const int w = wb + warp, t = w * 32 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, t < a.n_trials && flag[t] == 0);
            if (lane == 0 && w < n_words) s_mask[w] = m;
        }
    }
    __syncthreads();
    if (compact && !fused && tid == 0) {
        const int n_words = (a.n_trials + 31) >> 5;
        int acc = 0;
        for (int w = 0; w < n_words; ++w) { s_cnt[w] = acc; acc += __popc(s_mask[w]); }
        s_cnt[n_words] = acc;
    }
    const float4 *trial0 = a.trial + (long long)set * a.trial_set_stride;
    // staging: class by class, four 4-byte asynchronous copies per atom straight into the interleaved layout.  With the fused
    // screen the residue's near list goes first, as a copy group of its own
    if (fused) {
        const float4 *env = a.env + (long long)set * a.s.n_atoms;
        const int *slot = a.s.scr_slot + a.s.scr_off[a.step];
        float *l_flat = reinterpret_cast<float *>(s_list);
        for (int c = 0; c < n_cls; ++c) {
            const int len = s_lps[c + 1] - s_lps[c], real = s_lpre[c + 1] - s_lpre[c];
            for (int i = tid; i < len; i += SC_THREADS) {
                const int P = s_lps[c] + i;
                float *dst = l_flat + 8 * (P >> 1) + (P & 1);
                if (i < real) {
                    const float *src = reinterpret_cast<const float *>(env + slot[s_lpre[c] + i]);
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(src));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 8), "l"(src + 1));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 16), "l"(src + 2));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 24), "l"(src + 3));
                } else {
                    dst[0] = 1.0e15f; dst[2] = 1.0e15f; dst[4] = 1.0e15f; dst[6] = 0.f;
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    }
    {
        const float4 *env = a.env + (long long)set * a.s.n_atoms;
        for (int c = 0; c < n_cls; ++c) {
            const int first = s_first[c], len = s_ps[c + 1] - s_ps[c];            // (padded length)
            const int real = max(0, min(s_pre[c + 1], hi) - first);
            const int slot0 = a.s.cls_start[c] + (first - s_pre[c]);
            for (int i = tid; i < len; i += SC_THREADS) {
                const int P = s_ps[c] + i;
                float *dst = s_flat + 8 * (P >> 1) + (P & 1);
                if (i < real) {
                    const float *src = reinterpret_cast<const float *>(env + slot0 + i);
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(src));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 8), "l"(src + 1));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 16), "l"(src + 2));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 24), "l"(src + 3));
                } else {                                                          // pad slot of an odd segment: a parked dummy
                    dst[0] = 1.0e15f; dst[2] = 1.0e15f; dst[4] = 1.0e15f; dst[6] = 0.f;
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    }
    if (fused) {
        // ---- the neighbour screen, while the environment copies are in flight: every trial's chi-dependent atoms against the
        //      staged list, clash test only (clash_segment: the scoring pass's decision rule); the special-pair kernel's clash
        //      flags join; one flag per trial, then the compacted index of the survivors ----
        asm volatile("cp.async.wait_group 1;" ::: "memory");            // the list's group has landed (the environment's may not)
        __syncthreads();
        const int T = a.n_trials;
        const int tpb_s = max(1, SC_SE / n_per);
        for (int bt = 0; bt < T; bt += tpb_s) {
            const int nbt = min(tpb_s, T - bt);
            const int n_loc = nbt * n_per;
            for (int ch = warp; ch * 64 < n_loc; ch += nwarps) {
                float2 nx[SCORE_R], ny[SCORE_R], nz[SCORE_R];
                int ci[SCORE_R];
                bool clash[SCORE_R];
#pragma unroll
                for (int r = 0; r < SCORE_R; ++r) {
                    const int la = ch * 64 + r * 32 + lane;
                    int k = 0; float4 p = trial0[0];
                    if (la < n_loc) {
                        const int tl = la / n_per, kk = la - tl * n_per;
                        k = dedup ? st.var_k[kk] : kk;
                        p = trial0[(long long)(bt + tl) * n_sc + k];
                    }
                    nx[r] = make_float2(-p.x, -p.x); ny[r] = make_float2(-p.y, -p.y); nz[r] = make_float2(-p.z, -p.z);
                    ci[r] = a.s.cls_of_atom[st.sc_start + k];
                    clash[r] = false;
                }
                const bool two = ch * 64 + 32 < n_loc;
                for (int c = 0; c < n_cls; ++c) {
                    const int p0 = s_lps[c] >> 1, p1 = s_lps[c + 1] >> 1;
                    if (p0 >= p1) continue;
                    float thr2[SCORE_R];
#pragma unroll
                    for (int r = 0; r < SCORE_R; ++r) thr2[r] = __ldg(&a.s.cls_thr2f[ci[r] * n_cls + c]);
                    if (two) clash_segment<2>(s_list, p0, p1, nx, ny, nz, thr2, clash);
                    else clash_segment<1>(s_list, p0, p1, nx, ny, nz, thr2, clash);
                }
#pragma unroll
                for (int r = 0; r < SCORE_R; ++r) {
                    const int la = ch * 64 + r * 32 + lane;
                    if (la < n_loc) s_cl[la] = clash[r] ? 1 : 0;
                }
            }
            __syncthreads();
            for (int tl = tid; tl < nbt; tl += SC_THREADS) {
                int cl = 0;
                for (int k = 0; k < n_per; ++k) cl |= s_cl[tl * n_per + k];
                if (a.sp_partial[(long long)set * T + bt + tl].y != 0.0) cl = 1;      // clash among the special pairs
                s_flag[bt + tl] = (unsigned char)cl;
                a.screen[(long long)set * T + bt + tl] = (unsigned char)cl;           // (the selection reads the flags)
            }
            __syncthreads();
        }
        const int n_words = (T + 31) >> 5;
        for (int wb = 0; wb < n_words; wb += nwarps) {
            const int w = wb + warp, t = w * 32 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, t < T && s_flag[t] == 0);
            if (lane == 0 && w < n_words) s_mask[w] = m;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int w = 0; w < n_words; ++w) { s_cnt[w] = acc; acc += __popc(s_mask[w]); }
            s_cnt[n_words] = acc;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // the special environment atoms are scored by the special-pair kernel: parked here
    if (tid < st.n_sp_env) {
        const int slot = st.sp_env_slot[tid];
        int c_lo = 0, c_hi = n_cls;
        while (c_hi - c_lo > 1) { const int m = (c_lo + c_hi) >> 1; if (a.s.cls_start[m] <= slot) c_lo = m; else c_hi = m; }
        const int r = slot - a.s.cls_start[c_lo];
        if (r < s_pre[c_lo + 1] - s_pre[c_lo]) {
            const int ai = s_pre[c_lo] + r;
            if (ai >= lo && ai < hi) {
                const int P = s_ps[c_lo] + (ai - s_first[c_lo]);
                float *dst = s_flat + 8 * (P >> 1) + (P & 1);
                dst[0] = 1.0e15f; dst[2] = 1.0e15f; dst[4] = 1.0e15f; dst[6] = 0.f;
            }
        }
    }
    __syncthreads();
    const int n_keep = compact ? s_cnt[(a.n_trials + 31) >> 5] : a.n_trials;
    auto trial_of = [&](int j) -> int {
        if (!compact) return j;
        int w = 0;
        while (s_cnt[w + 1] <= j) ++w;
        return w * 32 + (int)__fns(s_mask[w], 0, j - s_cnt[w] + 1);
    };
    const int stride = a.n_trials + (dedup ? 1 : 0);
    const int n_fix = dedup ? st.n_fix : 0;
    const int tpb = max(1, ((fixed_shares > 0 ? min(SC_BATCH, SC_SE / fixed_shares) : SC_BATCH) - MAX_SC) / n_per);
    // (a conformer whose trials were all screened out still scores the chi-independent atoms: bt = 0 runs once)
    for (int bt = 0; bt < n_keep || (bt == 0 && n_fix > 0); bt += tpb) {
        const int nbt = max(0, min(tpb, n_keep - bt));
        const bool lastb = bt + tpb >= n_keep;
        const int n_varatoms = nbt * n_per;
        const int n_fixatoms = lastb ? n_fix : 0;
        const int n_local = n_varatoms + n_fixatoms;
        const int chunks = (n_local + 63) >> 6;
        int esplit = fixed_shares;
        if (fixed_shares <= 0) {
            esplit = 1;
            float best = 1.0e30f;
            const int e_max = min(8, SC_SE / max(1, n_local));
            for (int e = 1, units_e = chunks; e <= e_max; ++e, units_e += chunks) {
                const float cost = (float)((units_e + nwarps - 1) / nwarps) * (__frcp_rn((float)e) + 0.04f);
                if (cost < best) { best = cost; esplit = e; }
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
                int k = 0; float4 p = trial0[0];
                if (la < n_varatoms) {
                    const int tl = la / n_per, kk = la - tl * n_per;
                    k = dedup ? st.var_k[kk] : kk;
                    p = trial0[(long long)trial_of(bt + tl) * n_sc + k];
                } else if (la < n_local) {
                    k = st.fix_k[la - n_varatoms];
                    p = trial0[k];
                }
                katom[r] = k;
                nx[r] = make_float2(-p.x, -p.x); ny[r] = make_float2(-p.y, -p.y); nz[r] = make_float2(-p.z, -p.z);
                ci[r] = a.s.cls_of_atom[st.sc_start + k];
                e_lj[r] = 0.0; e_c[r] = 0.0; clash[r] = false;
            }
            if (ch * 64 + 32 < n_local) gc_score_unit<2>(s_env, s_ps, s_pp, a.s.cls_thr2f, a.s.cls_A, a.s.cls_B, n_cls, j, esplit, nx, ny, nz, ci, e_lj, e_c, clash);
            else gc_score_unit<1>(s_env, s_ps, s_pp, a.s.cls_thr2f, a.s.cls_A, a.s.cls_B, n_cls, j, esplit, nx, ny, nz, ci, e_lj, e_c, clash);
#pragma unroll
            for (int r = 0; r < SCORE_R; ++r) {
                const int la = ch * 64 + r * 32 + lane;
                if (la < n_local) {
                    const double qi = a.s.charge[st.sc_start + katom[r]];
                    s_e[j * n_local + la] = e_lj[r] + COULOMB_PREF * qi * e_c[r];
                    s_cl[j * n_local + la] = clash[r] ? 1 : 0;
                }
            }
        }
        __syncthreads();
        if (esplit > 1) {
            for (int la = tid; la < n_local; la += SC_THREADS) {
                double e = s_e[la]; int cl = s_cl[la];
                for (int j = 1; j < esplit; ++j) { e += s_e[j * n_local + la]; cl |= s_cl[j * n_local + la]; }
                s_e[la] = e; s_cl[la] = (unsigned char)cl;
            }
            __syncthreads();
        }
        for (int tl = tid; tl < nbt; tl += SC_THREADS) {
            double e = 0.0; int cl = 0;
            for (int k = 0; k < n_per; ++k) { e += s_e[tl * n_per + k]; cl |= s_cl[tl * n_per + k]; }
            a.partial[((long long)set * stride + trial_of(bt + tl)) * a.n_split + z] = make_double2(e, cl ? 1.0 : 0.0);
        }
        if (n_fixatoms > 0 && tid == SC_THREADS - 1) {
            double e = 0.0; int cl = 0;
            for (int k = 0; k < n_fixatoms; ++k) { e += s_e[n_varatoms + k]; cl |= s_cl[n_varatoms + k]; }
            a.partial[((long long)set * stride + a.n_trials) * a.n_split + z] = make_double2(e, cl ? 1.0 : 0.0);
        }
        __syncthreads();
    }
}
