// Neighbour-list pipeline (included at the end of kernels.cuh, inside namespace blobs).
//
// brute_force_collisions (physics.rs:241-317) tests every pair every substep; the grid pipeline (k_main + k_scan + k_scatter)
// re-sorts every collider into cells every substep to find the same pairs. But between two substeps a sphere moves by a small
// fraction of its radius, so WHO can touch whom changes slowly. Here every collider keeps a list of all colliders within
// r_a + r_b + skin of it, sorted by slot, and the per-substep kernel (k_step) only walks that list:
//   * the exact narrowphase (narrowphase(), the reference's arithmetic) runs on the CURRENT snapshots of the listed colliders,
//     so the contact set of every substep is exactly the reference's — the list only has to be a superset of it;
//   * the list is sorted by partner slot, which is the reference's pair-loop order for a single-collider body, so
//     contributions are simply added as they are found: no ordered insert, no contact list in local memory;
//   * every publisher tracks how far snapshots have moved since the lists were built (nl_track); k_nl_decide turns that into a
//     device-side "rebuild" flag, and the four rebuild kernels below (cell counting sort of the current snapshots + list
//     construction) return at once when it is not set. No host round trip, so whole steps stay CUDA-graph replayable.
// Bodies with more than NL_CAP neighbours (the shell the circle constraint builds, physics.rs:377-395) are handed to k_crowded,
// which walks the cell grid of the last rebuild instead of a list.

// ---------------------------------------------------------------------------------------------------------------------
// The decision: are the lists rebuilt before the coming substep's contact pass? Also predicts the common displacement c the
// publishers will subtract and resets the accumulators. Taken by the last CTA of k_step when that kernel is the substep's only
// publisher (SubstepParams::nl_tail_decide), else by the one-thread kernel k_nl_decide at the start of the substep; the first
// substep of every step call always runs k_nl_decide, which is also where a host request (NlCtl::force) is honoured.
// ---------------------------------------------------------------------------------------------------------------------
// (max_m_bits, sum_x, sum_y, n): the accumulators of the substep that just ended - this GPU's, or (strips) every rank's combined.
// xlim: strips only - also rebuild when the common displacement along x exceeds this (ownership is re-assigned at rebuilds only)
__device__ __forceinline__ void nl_decide_from(volatile NlCtl* ctl, unsigned int max_m_bits, float sum_x, float sum_y, unsigned int n, float lim,
                                               float xlim, uint32_t in_step) {
    const float m = __uint_as_float(max_m_bits);
    float mx = ctl->mean_x, my = ctl->mean_y;   // no sample this substep: assume nothing moved
    if (n) { mx = sum_x / (float)n; my = sum_y / (float)n; }
    if (!(fabsf(mx) < 1e30f)) mx = 0.f;
    if (!(fabsf(my) < 1e30f)) my = 0.f;
    const unsigned int need = (ctl->force != 0u || !(m <= lim) || !(fabsf(mx) <= xlim)) ? 1u : 0u;
    // displacement per substep (the references were reset by the last rebuild, hence mean_* == 0 right after one)
    const float ux = mx - ctl->mean_x, uy = my - ctl->mean_y;
    if (need) {   // references move to the current snapshots: the next publishers see one substep's worth of motion
        ctl->cx = ux; ctl->cy = uy;
        ctl->mean_x = 0.f; ctl->mean_y = 0.f;
        ctl->rebuilds += in_step;
    } else {      // linear extrapolation: under gravity alone the error is g * dt^2
        ctl->cx = mx + ux; ctl->cy = my + uy;
        ctl->mean_x = mx; ctl->mean_y = my;
    }
    ctl->need = need;
    ctl->force = 0u;
    ctl->max_m = 0u;
    ctl->sum_x = 0.f; ctl->sum_y = 0.f;
    ctl->n_sum = 0u;
    ctl->substeps += in_step;
}
__device__ __forceinline__ void nl_decide(volatile NlCtl* ctl, float lim, uint32_t in_step) {
    nl_decide_from(ctl, ctl->max_m, ctl->sum_x, ctl->sum_y, ctl->n_sum, lim, 3.4e38f, in_step);
}

// Inside a captured CUDA graph the rebuild kernels of a substep sit in the body of a conditional (IF) node: the kernel that takes
// the decision also hands it to that node (cudaGraphSetConditional), so a substep that keeps its lists launches nothing at all for
// the rebuild. With plain launches (no graph, profiling, the host-compiled test build) the same kernels are launched
// unconditionally and return at once when NlCtl::need is not set.
__device__ __forceinline__ void nl_set_cond(unsigned long long handle, unsigned int need) {
#ifndef BLOBS_EMU
    if (handle != 0ull) cudaGraphSetConditional((cudaGraphConditionalHandle)handle, need);
#endif
}

// the two tables are picked by ternaries (a runtime index into a kernel-parameter array would copy the struct to local memory)
__device__ __forceinline__ uint32_t* nl_tab(const NlView& L, uint32_t which) { return which ? L.tab[1] : L.tab[0]; }
__device__ __forceinline__ uint32_t* nl_tile(const NlView& L, uint32_t which) { return which ? L.tile[1] : L.tile[0]; }
constexpr unsigned NL_GATED_CTAS = 148 * 8;
// Which colliders a rebuild kernel visits: every slot - or, on a strip-decomposed world (arrays are indexed by GLOBAL slot there,
// 16 M slots for 2 M owned), the colliders of the bodies in the owned list. Returns NO_SLOT for "nothing at index i".
struct NlEnum {
    const uint32_t* olist;     // nullptr = all collider slots
    const uint32_t* ocount;
    const uint2* binfo;
    uint32_t n;                // collider slots (all-slot mode) / launch bound of the owned list
    __device__ __forceinline__ uint32_t limit() const { return olist != nullptr ? min(n, __ldg(ocount)) : n; }
    __device__ __forceinline__ uint32_t at(uint32_t i) const {
        if (olist == nullptr) return i;
        const uint32_t b = olist[i];
        if (b == NO_SLOT) return NO_SLOT;
        const int32_t col = (int32_t)binfo[b].y;
        return col >= 0 ? (uint32_t)col : NO_SLOT;
    }
};   // the gated kernels run grid-stride loops on a fixed grid: returning at once costs ~2 us

// ---------------------------------------------------------------------------------------------------------------------
// Rebuild, step 1-3: counting sort of the current snapshots into cells (the grid pipeline's k_count / k_scan / k_scatter, gated
// by the device flag and addressing the table pair by the device parity).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_nl_count(GridDesc g, ColliderArrays Cc, const uint32_t* __restrict__ bworld, NlView L, NlEnum E) {
    if (L.ctl->need == 0u) return;
    const uint32_t nx = (L.ctl->parity & 1u) ^ 1u;
    uint32_t* const tab_next = nl_tab(L, nx);
    uint32_t* const tile_next = nl_tile(L, nx);
    const uint32_t lim = E.limit();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < lim; i += gridDim.x * blockDim.x) {
        const uint32_t c = E.at(i);
        if (c == NO_SLOT) continue;
        if (!(Cc.cconst[c].y & CF_ACTIVE)) continue;
        const float2 a = Cc.cabs[c];
        const uint32_t wbase = g.n_worlds > 1u ? bworld[Cc.cparent[c]] * g.ncells : 0u;
        const uint32_t cell = wbase + cell_index(g, bin_coord(a.x, g.inv_cell), bin_coord(a.y, g.inv_cell));
        Cc.ccell[c] = make_uint2(cell, bin_collider(tab_next, tile_next, cell));
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_nl_scan(NlView L, uint32_t n) {
    if (L.ctl->need == 0u) return;
    const uint32_t cur = L.ctl->parity & 1u, nx = cur ^ 1u;
    // the table of the previous rebuild is zeroed on the way: it takes the counts of the NEXT rebuild
    scan_tile(nl_tab(L, nx), n, nl_tab(L, cur), n, nl_tile(L, nx), nl_tile(L, cur));
}

__global__ void __launch_bounds__(256) k_nl_scatter(ColliderArrays Cc, NlView L, NlEnum E) {
    if (L.ctl->need == 0u) return;
    const uint32_t* __restrict__ tab = nl_tab(L, (L.ctl->parity & 1u) ^ 1u);
    const uint32_t lim = E.limit();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < lim; i += gridDim.x * blockDim.x) {
        const uint32_t c = E.at(i);
        if (c == NO_SLOT) continue;
        const uint4 cc = Cc.cconst[c];
        if (!(cc.y & CF_ACTIVE)) continue;
        const uint2 cr = Cc.ccell[c];
        const float2 a = Cc.cabs[c];
        L.hot[__ldg(tab + cr.x) + cr.y] = make_float4(a.x, a.y, __uint_as_float(cc.x), __uint_as_float(hot_word(c, cc.y)));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Rebuild, step 4: one thread per collider slot walks the cells within r + r_max + skin of its snapshot (g.rmax is inflated by
// the skin on the host) and keeps every collider closer than r_a + r_b + skin, sorted by slot (sorted insert into a
// shared-memory column), in the transposed list array. Also (re)writes the slot-indexed snapshot record the next contact pass reads, and the
// reference position the displacement tracking measures from. The last CTA to finish flips the table parity.
// The keep test is deliberately loose (1e-4 relative): it only has to err on the side of keeping.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NL_BUILD_THREADS = 128;

__global__ void __launch_bounds__(NL_BUILD_THREADS) k_nl_build(GridDesc g, ColliderArrays Cc, const uint32_t* __restrict__ bworld, NlView L,
                                                               float4* __restrict__ snap_cur, NlEnum E, const uint8_t* __restrict__ cowned,
                                                               StripDesc S) {
    __shared__ uint32_t keep[NL_CAP * NL_BUILD_THREADS];
    NlCtl* const ctl = L.ctl;
    if (ctl->need == 0u) return;   // grid-uniform
    const uint32_t nx = (ctl->parity & 1u) ^ 1u;
    const uint32_t* __restrict__ tab = nl_tab(L, nx);
    const float4* __restrict__ hot = L.hot;
    const uint32_t tid = threadIdx.x;
    const uint32_t lim = E.limit();
    for (uint32_t i = blockIdx.x * (uint32_t)NL_BUILD_THREADS + tid; i < lim; i += gridDim.x * (uint32_t)NL_BUILD_THREADS) {
        const uint32_t c = E.at(i);
        if (c == NO_SLOT) continue;
        const uint4 cc = Cc.cconst[c];
        uint4 hd = make_uint4(0u, 0u, NL_INACTIVE, cc.y);
        if ((cc.y & CF_ACTIVE) && (cowned == nullptr || cowned[c])) {
            const float2 a = Cc.cabs[c];
            const float r = __uint_as_float(cc.x);
            snap_cur[c] = make_float4(a.x, a.y, r, __uint_as_float(hot_word(c, cc.y)));
            const uint32_t wbase = g.n_worlds > 1u ? bworld[Cc.cparent[c]] * g.ncells : 0u;
            const float rs = r + L.skin;
            uint32_t n = 0;
            auto offer = [&](const float4 h) {
                const uint32_t oslot = __float_as_uint(h.w) & HOT_SLOT_MASK;
                const float dx = a.x - h.x, dy = a.y - h.y;
                const float d2 = dx * dx + dy * dy;
                const float cut = (rs + h.z) * 1.0001f;
                if (oslot != c && !(d2 > cut * cut)) {   // NaNs are kept
                    if (n < (uint32_t)NL_CAP) {          // sorted insert into this thread's shared-memory column
                        uint32_t i = n;
                        while (i > 0u && keep[(i - 1u) * NL_BUILD_THREADS + tid] > oslot) {
                            keep[i * NL_BUILD_THREADS + tid] = keep[(i - 1u) * NL_BUILD_THREADS + tid];
                            --i;
                        }
                        keep[i * NL_BUILD_THREADS + tid] = oslot;
                    }
                    ++n;
                }
            };
            const CellRange R = cell_range(g, a.x, a.y, r);
            const uint32_t n1 = min(R.nx, g.W - R.c0);   // cells before the row wraps
            for (uint32_t j = 0; j < R.ny; ++j) {
                uint32_t row = R.r0 + j;
                if (row >= g.H) row -= g.H;
                const uint32_t base = wbase + row * g.W;
                uint32_t lo = __ldg(tab + base + R.c0), hi = __ldg(tab + base + R.c0 + n1);
#pragma unroll 4
                for (uint32_t k = lo; k < hi; ++k) offer(__ldg(hot + k));
                if (n1 < R.nx) {   // wrapped part of the row
                    lo = __ldg(tab + base);
                    hi = __ldg(tab + base + (R.nx - n1));
                    for (uint32_t k = lo; k < hi; ++k) offer(__ldg(hot + k));
                }
            }
            uint32_t cnt = NL_OVER;
            if (n <= (uint32_t)NL_CAP) {
                cnt = n;
                for (uint32_t i = 0; i < n; ++i) L.idx[(size_t)i * L.stride + c] = keep[i * NL_BUILD_THREADS + tid];
            }
            uint32_t fl = cc.y;
            if (cowned != nullptr) {   // strips: the neighbour that keeps this collider as a ghost (same test as strip_pack_one, which sent it there)
                const float reach = __fadd_ru(r, S.rmax);
                if (S.has_right && a.x >= __fsub_rd(S.x_hi, reach)) fl |= NLF_PUSH_R;
                if (S.has_left && a.x < __fadd_ru(S.x_lo, reach)) fl |= NLF_PUSH_L;
            }
            hd = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), cnt, fl);
        }
        L.hdr[c] = hd;
    }
    __syncthreads();
    if (tid == 0u) {
        __threadfence();
        if (atomicAdd(&ctl->done, 1u) == gridDim.x - 1u) {   // last CTA: every CTA has read the parity by now
            ctl->done = 0u;
            ctl->parity = nx;
        }
    }
}

__global__ void __launch_bounds__(32) k_nl_decide(NlCtl* ctl, float lim, uint32_t in_step, unsigned long long cond) {
    if (threadIdx.x != 0u || blockIdx.x != 0u) return;
    if (ctl->decided) {   // k_step's last CTA has decided already; only a host request can still change the verdict
        ctl->decided = 0u;
        if (ctl->force) {
            ctl->force = 0u;
            if (!ctl->need) {
                ctl->need = 1u;
                ctl->cx -= ctl->mean_x; ctl->cy -= ctl->mean_y;   // the references are about to move to the current snapshots
                ctl->mean_x = 0.f; ctl->mean_y = 0.f;
                ctl->rebuilds += in_step;
            }
        }
    } else {
        nl_decide(ctl, lim, in_step);
    }
    nl_set_cond(cond, ctl->need);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_step: one thread per body slot (bodies with zero or one collider; multi-collider bodies go to k_multi).
// brute_force_collisions restricted to the body's neighbour list + apply_gravity + update_objects + apply_constraints
// (physics.rs:241-395), then the new snapshot record and its displacement tracking.
// Memory rounds: (1) the body arrays, the own snapshot record, the list header and the first NL_SPEC list rows, all at the
// speculated collider slot c == b (lock-step insertion); (2) the listed snapshot records, NL_SPEC in flight, squared-distance
// prefilter into a bit mask; (3) survivors re-read (L1) one by one for the exact narrowphase, in list order = ascending slot =
// the reference's summation order.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool nl_prefilter(const SelfCol& s, float srk, const float4 h) {
    const float dx = s.x - h.x, dy = s.y - h.y;
    const float d2 = __fmaf_rn(dx, dx, dy * dy);
    const float mdk = __fmaf_rn(h.z, 1.00005f, srk);   // (ra + rb) * 1.00005, see gather_single
    return !(d2 > mdk * mdk);
}

// Rare path: a body whose collider has more neighbours than a list holds, met before the host has put k_crowded into the pipeline
// (that happens from the next call on). One thread does the WHOLE body - serial, exact contact pass over the cell grid of the last
// rebuild, then the same tail as k_step - out of line and called at the very end of k_step, where nothing else is live: an inlined
// or mid-kernel call costs every thread of k_step spills and a ~500-byte stack frame.
template <bool FUSED, bool STRIP>
__device__ BLOBS_NOINLINE void nl_over_body(SubstepParams P, GridDesc g, Constraints K, BodyArrays B, ColliderArrays Cc, Broadphase bp, Recording rec,
                                            DeviceStats* stats, uint32_t b) {
    const NlView& L = bp.nl;
    const uint2 info = B.binfo[b];
    const uint32_t flags = info.x;
    const uint32_t c = info.y;
    const float2 mg = B.bmg[b];
    float2 p = B.pos[b];
    const float2 po = B.pos_old[b];
    const float2 acc0 = load_acc(P, B, b);
    const bool hv = load_hv(P, B, b);
    const float4 me = L.snap_cur[c];
    const uint4 hd = reinterpret_cast<const uint4*>(L.hdr)[c];
    const uint32_t word = __float_as_uint(me.w);
    SelfCol s;
    s.x = me.x; s.y = me.y; s.r = me.z; s.m = mg.x;
    s.qx = __uint_as_float(hd.x); s.qy = __uint_as_float(hd.y);
    s.memb = s.filt = 0xffffffffu;
    if (word & HOT_COLD_BIT) {
        const uint4 cc = Cc.cconst[c];
        s.memb = cc.z; s.filt = cc.w;
    }
    s.body = b; s.slot = c; s.sensor = (word & HOT_SENSOR_BIT) != 0u;
    s.wbase = g.n_worlds > 1u ? B.bworld[b] * g.ncells : 0u;
    GatherOut out;
    out.fx = out.fy = 0.f;
    out.n_pairs = out.n_coinc = 0;
    const Broadphase gb = resolve_grid(bp);
    for_each_candidate(g, gb, Cc.ccold, s.wbase, s.qx, s.qy, s.r, [&](const Rec& o) {
        Contact ct;
        if (narrowphase(s, o, ct)) note_pair(s, o, ct, out, rec, B.vel, stats);
    });
    p = apply_contacts_rescan(g, gb, Cc.ccold, &s, 1, p.x, p.y);
    if (out.n_pairs) atomicAdd(&stats->collisions, (unsigned long long)out.n_pairs);
    if (out.n_coinc) atomicAdd(&stats->coincident, (unsigned long long)out.n_coinc);
    if (FUSED && !(flags & BF_JOINTED)) {
        float sx, sy, rot;
        integrate_body(P, K, B, b, flags, mg.y, p.x, p.y, po, acc0, hv, sx, sy, rot, stats);
        const float2 a = snapshot_of(Cc, c, hd.w, sx, sy, rot);
        Cc.cabs[c] = a;
        const float4 rec_new = make_float4(a.x, a.y, me.z, me.w);
        L.snap_next[c] = rec_new;
        if (STRIP) {
            if (hd.w & NLF_PUSH_L) L.peer_next[0][c] = rec_new;
            if (hd.w & NLF_PUSH_R) L.peer_next[1][c] = rec_new;
        }
        NlAcc na{0.f, 0.f, 0.f, 0u};
        nl_track(a.x, a.y, __uint_as_float(hd.x), __uint_as_float(hd.y), L.ctl->cx, L.ctl->cy, na);
        atomicMax(&L.ctl->max_m, __float_as_uint(na.m));   // (not part of the sampled mean displacement: an estimate anyway)
    } else {
        B.pos[b] = p;
    }
}

// End of a substep on a strip rank (one warp): this rank's displacement accumulators and a sequence number go to EVERY rank's flag
// block; the sequence number also tells the two neighbours that this substep's ghost records have landed in their arrays.
__device__ __forceinline__ void nls_publish_warp(NlCtl* ctl, const NlStripDev& X, uint32_t lane) {
    const volatile NlCtl* vc = ctl;
    const unsigned int seq = vc->pub_seq + 1u;
    const unsigned int m = vc->max_m, n = vc->n_sum;
    const float sx = vc->sum_x, sy = vc->sum_y;
    __syncwarp();
    if (lane < (uint32_t)X.nranks) {
        volatile NlFlag* f = X.peer[lane] + (size_t)(seq & 1u) * NL_MAX_RANKS + X.rank;
        f->max_m = m;
        f->sum_x = sx;
        f->sum_y = sy;
        f->n_sum = n;
        __threadfence_system();   // the numbers before the sequence number that announces them
        f->seq = seq;
    }
    __syncwarp();
    if (lane == 0u) {
        ctl->pub_seq = seq;
        ctl->published = 1u;
        ctl->max_m = 0u;
        ctl->sum_x = 0.f; ctl->sum_y = 0.f;
        ctl->n_sum = 0u;
    }
}

template <bool FUSED, int MINB, bool STRIP, int THREADS = 256>
__global__ void __launch_bounds__(THREADS, MINB) k_step(SubstepParams P, GridDesc g, Constraints K, BodyArrays B, ColliderArrays Cc, Broadphase bp,
                                                 Recording rec, DeviceStats* stats, NlStripDev X) {
    const NlView& L = bp.nl;
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    bool inb;
    if (STRIP) {   // strip-decomposed world: threads enumerate the compact list of bodies this rank owns
        inb = b < __ldg(L.ocount);
        b = inb ? L.olist[b] : NO_SLOT;
        inb = b != NO_SLOT;
    } else {
        inb = b < P.n_bodies;
    }
    const uint32_t bl = inb ? b : 0u;
    const float ccx = L.ctl->cx, ccy = L.ctl->cy;
    // round 1 (tail threads read slot 0 and discard)
    const uint2 info = B.binfo[bl];
    const float2 mg = B.bmg[bl];
    float2 p = B.pos[bl];
    const float2 po = B.pos_old[bl];
    const float2 acc0 = load_acc(P, B, bl);
    const bool hv = load_hv(P, B, bl);
    uint32_t cs = 0;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 hd = make_uint4(0u, 0u, NL_INACTIVE, 0u);
    uint32_t e[NL_SPEC];
#pragma unroll
    for (int k = 0; k < NL_SPEC; ++k) e[k] = 0u;
    if (P.n_colliders) {
        cs = min(bl, P.n_colliders - 1u);
        me = __ldg(L.snap_cur + cs);
        hd = __ldg(reinterpret_cast<const uint4*>(L.hdr) + cs);
#pragma unroll
        for (int k = 0; k < NL_SPEC; ++k) e[k] = __ldg(L.idx + (size_t)k * L.stride + cs);
    }
    const uint32_t flags = info.x;
    const int32_t col = (int32_t)info.y;
    NlAcc na{0.f, 0.f, 0.f, 0u};
    GatherOut out;
    out.fx = out.fy = 0.f;
    out.n_pairs = out.n_coinc = 0;
    unsigned int n_over = 0;
    bool over_self = false;
    if (inb && (flags & BF_ALIVE) && col >= BODY_NO_COLLIDER) {
        bool active_col = false, deferred = false;
        uint32_t c = 0;
        if (col >= 0) {
            c = (uint32_t)col;
            if (c != cs) {   // speculation missed: fetch the real collider
                me = __ldg(L.snap_cur + c);
                hd = __ldg(reinterpret_cast<const uint4*>(L.hdr) + c);
#pragma unroll
                for (int k = 0; k < NL_SPEC; ++k) e[k] = __ldg(L.idx + (size_t)k * L.stride + c);
            }
            active_col = hd.z != NL_INACTIVE;
            if (active_col && P.collisions_enabled) {
                const uint32_t word = __float_as_uint(me.w);
                SelfCol s;
                s.x = me.x; s.y = me.y; s.r = me.z; s.m = mg.x;
                s.qx = __uint_as_float(hd.x); s.qy = __uint_as_float(hd.y);   // where the grid of the last rebuild has this collider
                s.memb = s.filt = 0xffffffffu;
                if (word & HOT_COLD_BIT) {
                    const uint4 cc = Cc.cconst[c];
                    s.memb = cc.z; s.filt = cc.w;
                }
                s.body = b; s.slot = c; s.wbase = 0u; s.sensor = (word & HOT_SENSOR_BIT) != 0u;
                if (hd.z == NL_OVER) {   // more neighbours than a list holds (the shell the circle constraint builds)
                    n_over = 1;
                    if (P.crowded) {     // a whole warp of k_crowded does this body, pair counting included
                        P.over_list[atomicAdd(&stats->over_count[P.over_parity], 1u)] = OVER_COUNT_BIT | b;
                        deferred = true;
                    } else {             // first sighting (the host adds k_crowded to the pipeline from the next call on): this thread does
                        over_self = true;   // the whole body serially at the end of the kernel, where nothing else is live (nl_over_body)
                        deferred = true;
                    }
                } else {
                    const uint32_t cnt = hd.z;
                    const float srk = s.r * 1.00005f;
                    uint32_t mask = 0;
                    {
                        float4 h[NL_SPEC];
#pragma unroll
                        for (int k = 0; k < NL_SPEC; ++k) h[k] = __ldg(L.snap_cur + ((uint32_t)k < cnt ? e[k] : c));
#pragma unroll
                        for (int k = 0; k < NL_SPEC; ++k)
                            if ((uint32_t)k < cnt && nl_prefilter(s, srk, h[k])) mask |= 1u << k;
                    }
                    for (uint32_t k0 = NL_SPEC; k0 < cnt; k0 += 4u) {
                        uint32_t j[4];
                        float4 h[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) j[i] = k0 + i < cnt ? __ldg(L.idx + (size_t)(k0 + i) * L.stride + c) : c;
#pragma unroll
                        for (int i = 0; i < 4; ++i) h[i] = __ldg(L.snap_cur + j[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (k0 + i < cnt && nl_prefilter(s, srk, h[i])) mask |= 1u << (k0 + i);
                    }
                    while (mask) {   // the whole warp runs this loop together: trip count = most survivors of any lane
                        const uint32_t k = (uint32_t)__ffs(mask) - 1u;
                        mask &= mask - 1u;
                        const uint32_t j = __ldg(L.idx + (size_t)k * L.stride + c);
                        const Rec o = rec_of(__ldg(L.snap_cur + j), Cc.ccold);
                        Contact ct;
                        if (!narrowphase(s, o, ct)) continue;
                        note_pair(s, o, ct, out, rec, B.vel, stats);
                        if (ct.coincident) p.x = fadd(p.x, ct.i_am_a ? 0.01f : -0.01f);   // physics.rs:275-276, ordered just before the push
                        if (ct.push) { p.x = fadd(p.x, ct.cx); p.y = fadd(p.y, ct.cy); }
                    }
                }
            }
        }
        if (deferred) {
            // nothing: every array of this body is left untouched for k_crowded
        } else if (FUSED && !(flags & BF_JOINTED)) {   // jointed bodies are advanced by k_integrate after the joint projection
            float sx, sy, rot;
            integrate_body(P, K, B, b, flags, mg.y, p.x, p.y, po, acc0, hv, sx, sy, rot, stats);
            if (active_col) {
                const float2 a = snapshot_of(Cc, c, hd.w, sx, sy, rot);
                Cc.cabs[c] = a;
                const float4 rec_new = make_float4(a.x, a.y, me.z, me.w);
                L.snap_next[c] = rec_new;
                if (STRIP && (hd.w & (NLF_PUSH_L | NLF_PUSH_R))) {   // a neighbour rank keeps this collider as a ghost: store the record there too
                    if (hd.w & NLF_PUSH_L) L.peer_next[0][c] = rec_new;
                    if (hd.w & NLF_PUSH_R) L.peer_next[1][c] = rec_new;
                    // no fence here: the end-of-substep flag is written after a system-scope fence that follows the completion of
                    // every CTA of this kernel (its last CTA, or k_nls_publish) - one GPU-wide flush instead of one per pushing warp
                }
                nl_track(a.x, a.y, __uint_as_float(hd.x), __uint_as_float(hd.y), ccx, ccy, na);
            }
        } else {
            B.pos[b] = p;
        }
    }
    if (over_self) nl_over_body<FUSED, STRIP>(P, g, K, B, Cc, bp, rec, stats, b);
    nl_commit(L.ctl, na, &stats->collisions, out.n_pairs);
    if (P.nl_tail_decide && threadIdx.x == 0u) {   // this kernel is the substep's only publisher: the last CTA decides for the next substep
        __threadfence();
        if (atomicAdd(&L.ctl->step_done, 1u) == gridDim.x - 1u) {
            L.ctl->step_done = 0u;
            __threadfence();
            nl_decide(L.ctl, L.lim, 1u);
            L.ctl->decided = 1u;
            nl_set_cond(P.nl_cond_next, L.ctl->need);
        }
    }
    if (STRIP && P.nl_tail_publish && threadIdx.x < 32u) {   // strips, and this kernel is the substep's only publisher: the last CTA sends the flags
        unsigned int last = 0u;
        if (threadIdx.x == 0u) {
            __threadfence();   // this CTA's records and accumulator updates (ordered before thread 0 by nl_commit's barrier) before its arrival
            last = atomicAdd(&L.ctl->step_done, 1u) == gridDim.x - 1u ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (threadIdx.x == 0u) L.ctl->step_done = 0u;
            __threadfence_system();   // every CTA's peer stores before the flags that announce them
            nls_publish_warp(L.ctl, X, threadIdx.x);
        }
    }
    if (__any_sync(0xffffffffu, out.n_coinc | n_over)) {
        warp_add_u64(&stats->coincident, out.n_coinc);
        unsigned int o = __reduce_add_sync(0xffffffffu, n_over);
        if ((threadIdx.x & 31) == 0 && o) atomicAdd(&stats->list_overflow, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The list pipeline on a strip-decomposed world (BASELINE config #5; peer-memory exchange only).
// Between two rebuilds nothing structural crosses a strip edge: ownership is fixed, and so is the set of colliders each rank
// keeps as ghosts of its neighbours. k_step stores every new snapshot record of a collider flagged NLF_PUSH_L / _R straight into
// the neighbour's slot-indexed array (arrays are indexed by GLOBAL slot on every rank) - the "exchange" of a substep is those
// 16-byte peer stores plus one 32-byte flag per rank pair:
//   end of substep   k_nls_publish: this rank's displacement accumulators + a sequence number go to EVERY rank;
//   start of substep k_nls_decide: waits for every rank's flag of the previous substep (which also means their ghost records
//                    have landed), combines them in rank order and takes the rebuild decision - the same on every rank.
// A rebuild is collective: migrants change owner and the ghost sets are re-selected with the message exchange of the grid
// pipeline (k_strip_push), all of it gated by the device-side decision like the single-GPU rebuild kernels.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_nls_publish(NlCtl* ctl, NlStripDev X) {
    __threadfence_system();   // the previous kernels' peer stores are complete (stream order); flush them before the flags
    nls_publish_warp(ctl, X, threadIdx.x);
}

__global__ void __launch_bounds__(32) k_nls_decide(NlCtl* ctl, NlStripDev X, float lim, float xlim, void* send_l, void* send_r, uint32_t in_step,
                                                   DeviceStats* stats, unsigned long long cond) {
    const uint32_t lane = threadIdx.x;
    unsigned int M = 0u, N = 0u;
    float SX = 0.f, SY = 0.f;
    if (ctl->published) {
        const unsigned int seq = ctl->pub_seq;
        const volatile NlFlag* f = X.mine + (size_t)(seq & 1u) * NL_MAX_RANKS;
        unsigned int m = 0u, n = 0u;
        float sx = 0.f, sy = 0.f;
        if (lane < (uint32_t)X.nranks) {
            const long long t0 = clock64();
            while (f[lane].seq != seq) {
                if (clock64() - t0 > STRIP_WAIT_TICKS) { atomicOr(&stats->nan_flag, 8u); break; }
            }
            __threadfence_system();
            m = f[lane].max_m; sx = f[lane].sum_x; sy = f[lane].sum_y; n = f[lane].n_sum;
        }
        __syncwarp();
        for (int r = 0; r < X.nranks; ++r) {   // combined in rank order: every rank computes the very same numbers
            M = max(M, __shfl_sync(0xffffffffu, m, r));
            SX += __shfl_sync(0xffffffffu, sx, r);
            SY += __shfl_sync(0xffffffffu, sy, r);
            N += __shfl_sync(0xffffffffu, n, r);
        }
    }
    if (lane == 0u) {
        nl_decide_from(ctl, M, SX, SY, N, lim, xlim, in_step);
        if (ctl->need) {   // the rebuild packs migrants and ghost candidates into these messages
            StripHeader* hl = reinterpret_cast<StripHeader*>(send_l);
            StripHeader* hr = reinterpret_cast<StripHeader*>(send_r);
            hl->n_ghost = hl->n_mig = hl->overflow = 0u;
            hr->n_ghost = hr->n_mig = hr->overflow = 0u;
        }
        nl_set_cond(cond, ctl->need);
    }
}

// rebuild, strips: ghost candidates (and leavers) of all owned colliders into the two outgoing messages
__global__ void __launch_bounds__(256) k_nls_pack(BodyArrays B, ColliderArrays Cc, StripDesc S, NlEnum E, void* send_l, void* send_r, const NlCtl* ctl) {
    if (ctl->need == 0u) return;
    const uint32_t lim = E.limit();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < lim; i += gridDim.x * blockDim.x) {
        const uint32_t c = E.at(i);
        if (c == NO_SLOT) continue;
        const uint4 cc = Cc.cconst[c];
        if (!(cc.y & CF_ACTIVE)) continue;
        strip_pack_one(B, Cc, S, c, cc.y, Cc.cabs[c], __uint_as_float(cc.x), send_l, send_r);
    }
}

// which receive buffers the exchange with sequence number `seq` uses (k_strip_push layout: [from left / from right][parity])
__device__ __forceinline__ const void* nls_recv(const NlStripDev& X, int side, unsigned int seq) {
    return X.recv_block + ((size_t)side * 2u + (seq & 1u)) * X.stride;
}

__global__ void __launch_bounds__(256) k_nls_push(StripDesc S, const void* send_l, const void* send_r, NlStripDev X, const NlCtl* ctl,
                                                  DeviceStats* stats) {
    if (ctl->need == 0u) return;
    const uint32_t seq = *reinterpret_cast<volatile unsigned int*>(X.xseq) + 1u;
    const size_t par = seq & 1u;
    void* peer_l = S.has_left ? X.peer_block[0] + (1u * 2u + par) * X.stride : nullptr;    // I am my left neighbour's RIGHT
    void* peer_r = S.has_right ? X.peer_block[1] + (0u * 2u + par) * X.stride : nullptr;   // and my right neighbour's LEFT
    strip_push_body(S, send_l, send_r, peer_l, peer_r, nls_recv(X, 0, seq), nls_recv(X, 1, seq), seq, X.xseq, X.push_done, stats);
}

__global__ void __launch_bounds__(256) k_nls_bin_ghosts(GridDesc g, StripDesc S, NlStripDev X, NlView L, uint2* gcell, DeviceStats* stats) {
    if (L.ctl->need == 0u) return;
    const uint32_t nx = (L.ctl->parity & 1u) ^ 1u;
    const unsigned int seq = *X.xseq;   // k_nls_push has completed this exchange
    strip_bin_ghosts_body(g, S, nls_recv(X, 0, seq), nls_recv(X, 1, seq), nl_tab(L, nx), nl_tile(L, nx), gcell, stats);
}

__global__ void __launch_bounds__(256) k_nls_finish(BodyArrays B, ColliderArrays Cc, StripDesc S, const void* send_l, const void* send_r, NlStripDev X,
                                                    NlView L, const uint2* __restrict__ gcell, uint8_t* owned, uint8_t* cowned, uint32_t* olist,
                                                    uint32_t* ocount, uint32_t* opos, uint32_t olist_cap, DeviceStats* stats, float4* snap_cur) {
    if (L.ctl->need == 0u) return;
    const uint32_t nx = (L.ctl->parity & 1u) ^ 1u;
    const unsigned int seq = *X.xseq;
    strip_finish_body(B, Cc, S, send_l, send_r, nls_recv(X, 0, seq), nls_recv(X, 1, seq), nl_tab(L, nx), gcell, L.hot, owned, cowned, olist, ocount, opos,
                      olist_cap, stats, snap_cur);
}
