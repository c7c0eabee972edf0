// Device code of libblobs_b200: the per-substep kernels of Physics::integrate (reference
// blobs/src/physics.rs:397-422) for sm_100a.
//
// Arithmetic contract: every f32 operation on the parity path is written with the
// explicit round-to-nearest intrinsics (__fadd_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn). nvcc never
// contracts those into FMA, so evaluation order and rounding match Rust/glam (scalar IEEE f32, no
// FMA) irrespective of compiler flags. The TU is additionally built with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "types.cuh"

#ifdef BLOBS_EMU   // host-compiled test build: libstdc++ spells the attribute __noinline__ itself, so no macro of that name
#define BLOBS_NOINLINE __attribute__((noinline))
#else
#define BLOBS_NOINLINE __noinline__
#endif

namespace blobs {

// ------------------------------------------------------------------------------------------------
// exact f32 helpers (glam::Vec2 semantics: per-component scalar ops)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// glam Vec2::length = sqrt(x*x + y*y)
__device__ __forceinline__ float vlen(float x, float y) { return __fsqrt_rn(fadd(fmul(x, x), fmul(y, y))); }

// ------------------------------------------------------------------------------------------------
// cell arithmetic — SpatialHash::get_cell_coords (spatial.rs:57-62): floor(x / cs) as i32.
// __float2int_rd rounds toward -inf, saturates and maps NaN to 0 exactly like Rust's `as i32`.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(float v, float cs) { return __float2int_rd(fdiv(v, cs)); }
// Broadphase binning only needs a MONOTONE map shared by the binning and the query (see cell_range): floor(v * (1/cs)).
// (Identical to cell_coord whenever cs is a power of two, e.g. the cfg2 cell of 1.0.) The reference-exact
// get_cell_coords is what k_cell_coords exports.
__device__ __forceinline__ int bin_coord(float v, float inv_cs) { return __float2int_rd(v * inv_cs); }
// order-preserving int32 -> uint32, then modulo: a toroidal mapping where neighbouring cells stay neighbours
__device__ __forceinline__ uint32_t ubias(int c) { return (uint32_t)c ^ 0x80000000u; }
// Lemire fastmod: n % d for any 32-bit n, d with M = 2^64 / d + 1
__device__ __forceinline__ uint32_t fastmod(uint32_t n, unsigned long long M, uint32_t d) {
    return (uint32_t)__umul64hi(M * (unsigned long long)n, (unsigned long long)d);
}
__device__ __forceinline__ uint32_t cell_index(const GridDesc& g, int cx, int cy) {
    return fastmod(ubias(cy), g.MH, g.H) * g.W + fastmod(ubias(cx), g.MW, g.W);
}

__device__ __forceinline__ Rec make_rec(const float4 h, const uint4 c) {
    Rec r;
    r.x = h.x; r.y = h.y; r.r = h.z; r.slot_sensor = __float_as_uint(h.w);
    r.m = __uint_as_float(c.x); r.memb = c.y; r.filt = c.z; r.parent = c.w;
    return r;
}
__device__ __forceinline__ Rec load_rec(const Broadphase& bp, const uint4* __restrict__ ccold, uint32_t k) {
    float4 h = __ldg(bp.hot + k);
    // list pipeline: the cell-sorted array dates from the last list rebuild - it still says WHO is near (within the skin), the
    // record itself is re-read from the slot-indexed snapshot array of the previous substep
    if (bp.snap != nullptr) h = __ldg(bp.snap + (__float_as_uint(h.w) & HOT_SLOT_MASK));
    const uint32_t w = __float_as_uint(h.w);
    if (w & HOT_COLD_BIT) return make_rec(h, __ldg(ccold + (w & HOT_SLOT_MASK)));
    // default sphere: calculated_mass = 2 * (2r) exactly, groups ALL; its parent has no other collider, so it can never equal
    // the querying body (NO_SLOT is a safe stand-in; event recording, the only consumer of the real parent, forces the flag)
    return make_rec(h, make_uint4(__float_as_uint(fmul(4.0f, h.z)), 0xffffffffu, 0xffffffffu, NO_SLOT));
}

// Cell range that can hold a partner of a sphere at (x, y) with radius r, on the toroidal table.
// Coverage proof: a contact needs fl(dist) < fl(ra+rb), which implies |xa-xb| < r + rmax in
// real arithmetic; the range is computed from directed-rounding bounds of x -/+ (r + rmax) and cell_coord is monotonic.
struct CellRange {
    uint32_t c0, nx;   // first column (wrapped), number of columns (<= W)
    uint32_t r0, ny;   // first row (wrapped), number of rows (<= H)
};
__device__ __forceinline__ CellRange cell_range(const GridDesc& g, float x, float y, float r) {
    const float reach = __fadd_ru(r, g.rmax);
    const int cx0 = bin_coord(__fsub_rd(x, reach), g.inv_cell), cx1 = bin_coord(__fadd_ru(x, reach), g.inv_cell);
    const int cy0 = bin_coord(__fsub_rd(y, reach), g.inv_cell), cy1 = bin_coord(__fadd_ru(y, reach), g.inv_cell);
    CellRange R;
    // spans (>= 1); an empty/NaN range degenerates to one cell
    R.nx = (cx1 >= cx0) ? (uint32_t)cx1 - (uint32_t)cx0 + 1u : 1u;
    R.ny = (cy1 >= cy0) ? (uint32_t)cy1 - (uint32_t)cy0 + 1u : 1u;
    R.c0 = fastmod(ubias(cx0), g.MW, g.W);
    if (R.nx == 0u || R.nx >= g.W) { R.nx = g.W; R.c0 = 0; }
    R.r0 = fastmod(ubias(cy0), g.MH, g.H);
    if (R.ny == 0u || R.ny >= g.H) { R.ny = g.H; R.r0 = 0; }
    return R;
}

// Generic neighbourhood walk: calls f(const Rec&) for every record in the range (any span, row/column wrap).
template <class F>
__device__ __forceinline__ void for_each_candidate(const GridDesc& g, const Broadphase& bp, const uint4* __restrict__ ccold, uint32_t wbase,
                                                   float x, float y, float r, F&& f) {
    const CellRange R = cell_range(g, x, y, r);
    const uint32_t n1 = min(R.nx, g.W - R.c0);  // cells before the row wraps
    for (uint32_t j = 0; j < R.ny; ++j) {
        uint32_t row = R.r0 + j;
        if (row >= g.H) row -= g.H;
        const uint32_t base = wbase + row * g.W;
        uint32_t lo = __ldg(bp.tab + base + R.c0), hi = __ldg(bp.tab + base + R.c0 + n1);
        for (uint32_t k = lo; k < hi; ++k) f(load_rec(bp, ccold, k));
        if (n1 < R.nx) {  // wrapped part of the row
            lo = __ldg(bp.tab + base);
            hi = __ldg(bp.tab + base + (R.nx - n1));
            for (uint32_t k = lo; k < hi; ++k) f(load_rec(bp, ccold, k));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// narrowphase for one (self collider, candidate record). Restates the pair-loop body of
// brute_force_collisions (physics.rs:250-312) from the point of view of ONE of the two bodies.
// Returns false if there is no contact. On contact: key orders the contribution inside the reference's
// (i, j<i) loop; (cx, cy) is what the reference adds to THIS body's position.
// ------------------------------------------------------------------------------------------------
struct SelfCol {
    float x, y, r, m;
    float qx, qy;    // centre of the cell-range query: (x, y), or - list pipeline - where the collider was when the grid was built
    uint32_t memb, filt;
    uint32_t body;   // parent body slot
    uint32_t slot;   // collider slot
    uint32_t wbase;  // first table entry of this body's (batched) world
    bool sensor;
};

struct Contact {
    float cx, cy;
    bool push;         // false for sensor pairs (physics.rs:291)
    bool coincident;   // distance < 1e-6 branch (physics.rs:272-286)
    bool i_am_a;       // this collider is `col_a` (the later slot)
    uint32_t other;    // partner collider slot
};

__device__ __forceinline__ bool narrowphase(const SelfCol& s, const Rec& o, Contact& c) {
    if ((o.slot_sensor & HOT_SLOT_MASK) == s.slot) return false;                    // a collider never pairs with itself (j < i)
    if (o.parent == s.body) return false;                                           // physics.rs:260 (same parent)
    if (!((s.memb & o.filt) != 0u && (o.memb & s.filt) != 0u)) return false;        // groups.rs:52-57
    const uint32_t oslot = o.slot_sensor & HOT_SLOT_MASK;
    const bool osens = (o.slot_sensor & HOT_SENSOR_BIT) != 0u;
    const bool i_am_a = s.slot > oslot;                                             // physics.rs:248-249 (a = later slot)
    // axis = abs_a - abs_b (physics.rs:264); x - y == -(y - x) exactly, so compute from self and flip
    float ax = i_am_a ? fsub(s.x, o.x) : fsub(o.x, s.x);
    float ay = i_am_a ? fsub(s.y, o.y) : fsub(o.y, s.y);
    float dist = vlen(ax, ay);
    const float min_dist = i_am_a ? fadd(s.r, o.r) : fadd(o.r, s.r);                // physics.rs:267
    if (!(dist < min_dist)) return false;                                           // physics.rs:269
    c.coincident = false;
    c.i_am_a = i_am_a;
    c.other = oslot;
    if (dist < 1e-6f) {                                                             // physics.rs:272-286
        // Jacobi restatement of the push-out: a moves +0.01 x, b moves -0.01 x, snapshots follow.
        // (The reference reads live positions here; exact only when neither body was touched earlier
        // in the same pass — counted in stats.coincident_pairs, see DESIGN.md.)
        c.coincident = true;
        const float xa = i_am_a ? s.x : o.x, xb = i_am_a ? o.x : s.x;
        ax = fsub(fadd(xa, 0.01f), fsub(xb, 0.01f));
        dist = vlen(ax, ay);
    }
    c.push = !(s.sensor || osens);                                                  // physics.rs:291
    c.cx = 0.f;
    c.cy = 0.f;
    if (c.push) {
        const float nx = fdiv(ax, dist), ny = fdiv(ay, dist);                       // physics.rs:292
        const float delta = fsub(min_dist, dist);                                   // physics.rs:294
        const float ma = i_am_a ? s.m : o.m, mb = i_am_a ? o.m : s.m;
        const float ratio = fsub(1.0f, fdiv(ma, fadd(ma, mb)));                     // physics.rs:319-321
        if (i_am_a) {                                                               // physics.rs:298
            const float k = fmul(ratio, delta);
            c.cx = fmul(k, nx);
            c.cy = fmul(k, ny);
        } else {                                                                    // physics.rs:299 (p -= v == p += -v)
            const float k = fmul(fsub(1.0f, ratio), delta);
            c.cx = -fmul(k, nx);
            c.cy = -fmul(k, ny);
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Ordered contact list. The reference adds contributions to a body in pair-loop order:
// lexicographic (later slot, earlier slot). KEY = uint64 (later << 33 | earlier << 1 | bit) for bodies with several
// colliders; for a single-collider body the order degenerates to ascending partner slot, so a uint32 key
// (partner << 1 | bit) is enough. bit 0 orders the coincident push-out (physics.rs:275-276) just before the contact
// push of the same pair.
// ------------------------------------------------------------------------------------------------
constexpr int LIST_CAP = 24;

template <class KEY>
struct ContactList {
    KEY key[LIST_CAP];
    float cx[LIST_CAP], cy[LIST_CAP];
    int n;
    bool overflow;
    __device__ __forceinline__ void clear() { n = 0; overflow = false; }
    __device__ __forceinline__ void insert(KEY k, float x, float y) {
        if (n == LIST_CAP) { overflow = true; return; }
        int i = n++;
        while (i > 0 && key[i - 1] > k) {
            key[i] = key[i - 1]; cx[i] = cx[i - 1]; cy[i] = cy[i - 1];
            --i;
        }
        key[i] = k; cx[i] = x; cy[i] = y;
    }
};

template <class KEY>
__device__ __forceinline__ KEY pair_key(uint32_t s, uint32_t o, bool coincident);
template <>
__device__ __forceinline__ unsigned long long pair_key<unsigned long long>(uint32_t s, uint32_t o, bool coincident) {
    const uint32_t hi = s > o ? s : o, lo = s > o ? o : s;
    return ((unsigned long long)hi << 33) | ((unsigned long long)lo << 1) | (coincident ? 0ull : 1ull);
}
template <>
__device__ __forceinline__ uint32_t pair_key<uint32_t>(uint32_t, uint32_t o, bool coincident) {
    return (o << 1) | (coincident ? 0u : 1u);
}

__device__ __forceinline__ void warp_add_u64(unsigned long long* dst, unsigned int v) {
    // one atomic per warp; all 32 lanes must call
    unsigned int s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst, (unsigned long long)s);
}

struct GatherOut {
    float fx, fy;              // fast-mode running sum
    unsigned int n_pairs, n_coinc;
};

// Bookkeeping for one contact: the later slot reports the pair, once, as (a, b) (physics.rs:302-311).
__device__ __forceinline__ void note_pair(const SelfCol& s, const Rec& o, const Contact& c, GatherOut& out, const Recording& rec,
                                          const float2* __restrict__ vel, DeviceStats* stats) {
    if (!c.i_am_a) return;
    out.n_pairs++;
    if (c.coincident) out.n_coinc++;
    if (rec.mode) {
        const unsigned long long idx = atomicAdd(rec.count, 1ull);
        if (idx < rec.cap) {
            rec.pairs[idx] = make_uint2(s.slot, c.other);
            if (rec.mode == 2u) {
                // calculated_velocity before this substep's update (physics.rs:288-289); event mode always runs the
                // split pipeline, so vel[] is not being rewritten concurrently
                const float2 va = vel[s.body], vb = vel[o.parent];  // event mode flags every record needs_cold, so parent is real
                rec.vels[idx] = make_float4(va.x, va.y, vb.x, vb.y);
            }
        } else {
            atomicAdd(&stats->rec_dropped, 1ull);
        }
    }
}

// One confirmed candidate: narrowphase, bookkeeping, ordered insert (or fast-mode sum).
template <bool ORDERED, class KEY>
__device__ __forceinline__ void take_candidate(const SelfCol& s, const Rec& o, ContactList<KEY>& list, GatherOut& out,
                                               const Recording& rec, const float2* __restrict__ vel, DeviceStats* stats) {
    Contact c;
    if (!narrowphase(s, o, c)) return;
    note_pair(s, o, c, out, rec, vel, stats);
    if (c.coincident) {
        const float px = c.i_am_a ? 0.01f : -0.01f;
        if (ORDERED) list.insert(pair_key<KEY>(s.slot, c.other, true), px, 0.0f);
        else out.fx = fadd(out.fx, px);
    }
    if (c.push) {
        if (ORDERED) list.insert(pair_key<KEY>(s.slot, c.other, false), c.cx, c.cy);
        else { out.fx = fadd(out.fx, c.cx); out.fy = fadd(out.fy, c.cy); }
    }
}

// Generic gather (any cell span): used by k_multi and as the slow path of gather_single.
template <bool ORDERED, class KEY>
__device__ __forceinline__ void gather_generic(const GridDesc& g, const Broadphase& bp, const uint4* __restrict__ ccold, const SelfCol& s,
                                               ContactList<KEY>& list, GatherOut& out, const Recording& rec, const float2* __restrict__ vel,
                                               DeviceStats* stats) {
    for_each_candidate(g, bp, ccold, s.wbase, s.qx, s.qy, s.r, [&](const Rec& o) { take_candidate<ORDERED, KEY>(s, o, list, out, rec, vel, stats); });
}

// Latency- and divergence-oriented gather for the common case (<= 3 rows, no column wrap):
//  1. the six cell-table reads are issued together;
//  2. SCAN: the three row spans are treated as one flattened candidate sequence; hot halves (16 B: x, y, r, slot) are
//     fetched BATCH at a time (all loads in flight before the first use) and a conservative squared-distance
//     prefilter marks survivors in a 32-bit mask — no sqrt, no divide, no cold half;
//  3. RESOLVE: survivors are popped in a loop that all lanes of the warp run together (trip count = max survivors per
//     lane), each doing the exact narrowphase on hot + cold halves.
// Prefilter soundness: a contact needs fl(sqrt(d2)) < md (md = fl(ra+rb)); d2 > (md*1.00005)^2 implies sqrt(d2) > md*(1+4e-5)
// even after the ~1e-7 relative rounding of the (FMA-contracted, order-free) prefilter arithmetic, which rounding of the
// exact path (2^-24) cannot bring below md. NaNs fail the '>' and fall through to the exact test.
template <bool ORDERED, class KEY, int BATCH>
__device__ __forceinline__ void gather_single(const GridDesc& g, const Broadphase& bp, const uint4* __restrict__ ccold, const SelfCol& s,
                                              ContactList<KEY>& list, GatherOut& out, const Recording& rec, const float2* __restrict__ vel,
                                              DeviceStats* stats) {
    const CellRange R = cell_range(g, s.x, s.y, s.r);
    if (R.ny > 3u || R.c0 + R.nx > g.W) {
        gather_generic<ORDERED, KEY>(g, bp, ccold, s, list, out, rec, vel, stats);
        return;
    }
    uint32_t lo[3], cnt[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint32_t row = R.r0 + j;
        if (row >= g.H) row -= g.H;
        const bool valid = (uint32_t)j < R.ny;
        const uint32_t idx = s.wbase + (valid ? row * g.W + R.c0 : 0u);
        const uint32_t a = __ldg(bp.tab + idx), b = __ldg(bp.tab + idx + (valid ? R.nx : 0u));
        lo[j] = a;
        cnt[j] = b - a;
    }
    const uint32_t n0 = cnt[0], n01 = cnt[0] + cnt[1], total = n01 + cnt[2];
    // flattened candidate index -> record index; rebased so that k = t + off_j inside row j
    const uint32_t off0 = lo[0], off1 = lo[1] - n0, off2 = lo[2] - n01;
    const float srk = s.r * 1.00005f;
    for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t lim = min(32u, total - base);
        uint32_t mask = 0;
        for (uint32_t t0 = 0; t0 < lim; t0 += BATCH) {
            float4 h[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const uint32_t t = base + t0 + i;
                const uint32_t k = t + (t < n0 ? off0 : (t < n01 ? off1 : off2));
                h[i] = __ldg(bp.hot + (t0 + i < lim ? k : lo[0]));  // out-of-range slots re-read a valid record (arrays are padded by 1)
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const uint32_t oslot = __float_as_uint(h[i].w) & HOT_SLOT_MASK;
                const float dx = s.x - h[i].x, dy = s.y - h[i].y;
                const float d2 = __fmaf_rn(dx, dx, dy * dy);
                const float mdk = __fmaf_rn(h[i].z, 1.00005f, srk);  // (ra + rb) * 1.00005
                if (t0 + i < lim && oslot != s.slot && !(d2 > mdk * mdk)) mask |= 1u << (t0 + i);
            }
        }
        while (mask) {
            const uint32_t t = base + (uint32_t)__ffs(mask) - 1u;
            mask &= mask - 1u;
            const uint32_t k = t + (t < n0 ? off0 : (t < n01 ? off1 : off2));
            take_candidate<ORDERED, KEY>(s, load_rec(bp, ccold, k), list, out, rec, vel, stats);
        }
    }
}

// Rare path when a body has more than LIST_CAP contributions: windowed selection. Every pass re-scans the neighbourhood and
// keeps the LIST_CAP smallest keys greater than the last one applied (bounded insertion), applies them in order and moves
// the window on: ceil(k / LIST_CAP) scans for k contributions, same summation order as the reference. (The first version
// selected ONE key per scan; in the dense shell that forms where the circle constraint projects bodies onto its boundary a
// few bodies with hundreds of contacts then stalled the whole launch.) Everything by value so callers keep their state in
// registers.
__device__ __forceinline__ float2 apply_contacts_rescan(GridDesc g, Broadphase bp, const uint4* __restrict__ ccold, const SelfCol* cols, int ncols,
                                                        float px, float py) {
    bool have_last = false;
    unsigned long long last = 0;
    for (;;) {
        ContactList<unsigned long long> win;
        win.clear();
        auto offer = [&](unsigned long long k, float x, float y) {
            if (have_last && !(k > last)) return;
            if (win.n == LIST_CAP) {
                if (!(k < win.key[LIST_CAP - 1])) return;
                win.n = LIST_CAP - 1;  // drop the largest to make room
            }
            win.insert(k, x, y);
        };
        for (int ci = 0; ci < ncols; ++ci) {
            const SelfCol s = cols[ci];
            for_each_candidate(g, bp, ccold, s.wbase, s.qx, s.qy, s.r, [&](const Rec& o) {
                Contact c;
                if (!narrowphase(s, o, c)) return;
                if (c.coincident) offer(pair_key<unsigned long long>(s.slot, c.other, true), c.i_am_a ? 0.01f : -0.01f, 0.f);
                if (c.push) offer(pair_key<unsigned long long>(s.slot, c.other, false), c.cx, c.cy);
            });
        }
        if (win.n == 0) break;
        for (int i = 0; i < win.n; ++i) { px = fadd(px, win.cx[i]); py = fadd(py, win.cy[i]); }
        last = win.key[win.n - 1];
        have_last = true;
        if (win.n < LIST_CAP) break;
    }
    return make_float2(px, py);
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative contact resolution (k_main<POOLED>), for agitated / contact-rich states. Measured on config #2 once the
// block has hit the circle constraint (profiles/r2_dense_ncu.md): the reference's positional solver turns the pile into a hot
// gas - bodies move several diameters per substep, so slot order says nothing about position any more - with clusters where a
// body has 60+ candidates next to bodies with 5. One thread per body then runs at the pace of the busiest lane: 8 of 32 lanes
// active in the candidate scan, 12 of 32 overall, and the ordered insert into a per-thread list lives in local memory.
// Here the warp shares ALL the work of its 32 bodies:
//   1. every lane looks up its three row spans (six table loads) and publishes its collider + spans in shared memory;
//   2. the candidates of all lanes are flattened into one sequence (prefix sum of the span lengths) and scanned 64 per
//      iteration, one candidate per lane and step: owner found by binary search in the prefix array, squared-distance prefilter
//      (no sqrt / divide), survivors appended to a per-warp queue in sequence order (ballot + popc), which keeps every owner's
//      survivors contiguous;
//   3. exact narrowphase, one (owner, candidate) PAIR per lane; pair counting / recording happens here;
//   4. ordering = RANK of each contribution among its owner's (keys are unique), written as a permutation;
//   5. each owner adds its contributions in rank order = ascending partner slot = the reference's pair-loop order.
// Same arithmetic, same summation order as the per-lane path => bit-identical. Lanes are batched so that the candidates of a
// batch fit the queue (COOP_Q); a body with more candidates than that (the shell the circle constraint builds) goes to
// k_crowded, or - when that kernel is not in the pipeline - through the per-lane path. An owner that meets a coincident pair
// (distance < 1e-6: two contributions per pair, physics.rs:272-286) redoes its sum with the serial windowed rescan.
// ------------------------------------------------------------------------------------------------
constexpr int COOP_Q = 256;            // survivor queue entries per warp
constexpr uint32_t COOP_C = 1024u;     // candidates scanned per batch of lanes (if their survivors overflow the queue the batch is halved)
constexpr uint32_t COOP_LANE_MAX = 96u;   // a lane with more candidates than this is not pooled (rank ordering is quadratic in the run length)
constexpr uint32_t POOL_NONE = 0xffffffffu;

struct CoopSmem {                 // per warp
    float4 sa[32];                // the lanes' own colliders: x, y, r, parent mass
    uint4 sb[32];                 //                           memberships, filter, body slot, collider slot | sensor << 31
    uint32_t lo[3][32];           // first record of each row span
    uint16_t c1[32], c2[32];      // candidates in span 0, in spans 0 + 1 (pooled lanes: <= COOP_LANE_MAX)
    uint32_t pre[32];             // batch-relative exclusive prefix of the lanes' candidate totals
    uint32_t ol[32];              // the lanes of the batch that have candidates, in lane order
    uint32_t key[COOP_Q];         // survivor: record index, then the contribution key (POOL_NONE = contributes nothing)
    float cx[COOP_Q], cy[COOP_Q];
    uint16_t perm[COOP_Q];        // perm[segment start + rank] = entry
    uint8_t owner[COOP_Q];        // lane that owns the entry
    uint32_t seg[32], send[32];   // first / one-past-last queue entry of each owner lane
    uint32_t nvalid[32];          // contributions per owner lane
    uint32_t fb;                  // owner lanes that must fall back to the serial rescan
};

__device__ __forceinline__ Rec rec_of(const float4 h, const uint4* __restrict__ ccold) {
    const uint32_t w = __float_as_uint(h.w);
    if (w & HOT_COLD_BIT) return make_rec(h, __ldg(ccold + (w & HOT_SLOT_MASK)));
    return make_rec(h, make_uint4(__float_as_uint(fmul(4.0f, h.z)), 0xffffffffu, 0xffffffffu, NO_SLOT));   // default sphere, see load_rec
}

// Must be called by all 32 lanes of the warp (valid = this lane has a collider to resolve). Returns true when the lane's
// contributions were added to (px, py) here; false when they are in `list` (per-lane path), to be applied by the caller -
// unless `big` comes back set (only with defer_big): then nothing was done for this lane, not even pair counting.
template <int BATCH>
__device__ __forceinline__ bool gather_coop(const GridDesc& g, const Broadphase& bp, const uint4* __restrict__ ccold, bool valid, const SelfCol& s,
                                            ContactList<uint32_t>& list, GatherOut& out, const Recording& rec, const float2* __restrict__ vel,
                                            DeviceStats* stats, CoopSmem& ps, bool defer_big, bool& big, float& px, float& py) {
    constexpr uint32_t FULL = 0xffffffffu;
    big = false;
    const uint32_t lane = threadIdx.x & 31u;
    // ---- 1. spans -------------------------------------------------------------------------------------------------
    bool plain = false;
    uint32_t lo0 = 0, lo1 = 0, lo2 = 0, n0 = 0, n01 = 0, total = 0;
    if (valid) {
        const CellRange R = cell_range(g, s.x, s.y, s.r);
        if (!(R.ny > 3u || R.c0 + R.nx > g.W)) {
            uint32_t lo[3], cnt[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint32_t row = R.r0 + j;
                if (row >= g.H) row -= g.H;
                const bool rv = (uint32_t)j < R.ny;
                const uint32_t idx = s.wbase + (rv ? row * g.W + R.c0 : 0u);
                const uint32_t a = __ldg(bp.tab + idx), b = __ldg(bp.tab + idx + (rv ? R.nx : 0u));
                lo[j] = a;
                cnt[j] = b - a;
            }
            lo0 = lo[0]; lo1 = lo[1]; lo2 = lo[2];
            n0 = cnt[0]; n01 = cnt[0] + cnt[1]; total = n01 + cnt[2];
            plain = true;
        }
    }
    const bool coop = plain && total <= COOP_LANE_MAX;
    const uint32_t tc = coop ? total : 0u;
    ps.sa[lane] = make_float4(s.x, s.y, s.r, s.m);
    ps.sb[lane] = make_uint4(s.memb, s.filt, s.body, s.slot | (s.sensor ? HOT_SENSOR_BIT : 0u));
    ps.lo[0][lane] = lo0; ps.lo[1][lane] = lo1; ps.lo[2][lane] = lo2;
    ps.c1[lane] = (uint16_t)min(n0, 0xffffu); ps.c2[lane] = (uint16_t)min(n01, 0xffffu);
    ps.nvalid[lane] = 0u;
    if (lane == 0) ps.fb = 0u;
    uint32_t incl = tc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t excl = incl - tc;
    const uint32_t le_mask = FULL >> (31u - lane);   // bits 0 .. lane
    bool applied = false;
    uint32_t start = 0;
    while (start < 32u) {   // lane ranges [start, end): at most COOP_C candidates, and survivors that fit the queue
        const uint32_t base = __shfl_sync(FULL, excl, start);
        uint32_t end;
        {
            const bool in_batch = lane >= start && (incl - base) <= COOP_C;
            const uint32_t zeros = ~__ballot_sync(FULL, in_batch) & (FULL << start);
            end = zeros ? (uint32_t)__ffs(zeros) - 1u : 32u;   // > start: one pooled lane never exceeds COOP_LANE_MAX <= COOP_C
        }
        uint32_t qn, C;
        for (;;) {
            C = __shfl_sync(FULL, incl, end - 1u) - base;      // candidates of this batch
            const bool inb = lane >= start && lane < end;
            const uint32_t have = __ballot_sync(FULL, inb && tc != 0u);
            __syncwarp();
            ps.pre[lane] = excl - base;
            if (inb && tc != 0u) ps.ol[__popc(have & (le_mask >> 1))] = lane;
            __syncwarp();
            // ---- 2. cooperative scan. Candidate i of the batch belongs to the lane l with pre[l] <= i < pre[l] + total[l]. The 32
            // candidates of one group belong to consecutive owners: every owner marks where it starts inside the group (one OR
            // across the warp), so the owner of candidate t is the (number of marks at or before t)-th one, counted from `carry`.
            qn = 0;                 // warp-uniform queue length
            bool ovf = false;       // warp-uniform
            uint32_t carry = 0;     // owners that start before the current group
            for (uint32_t i0 = 0; i0 < C; i0 += 32u * BATCH) {
                float4 h[BATCH];
                uint32_t kk[BATCH], own[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const uint32_t I = i0 + 32u * u, i = I + lane;
                    const uint32_t myp = excl - base;   // (meaningful for the owners of this batch only)
                    const uint32_t M = __reduce_or_sync(FULL, (inb && tc != 0u && myp >= I && myp < I + 32u) ? 1u << (myp - I) : 0u);
                    const uint32_t ord = carry + (uint32_t)__popc(M & le_mask) - 1u;
                    carry += (uint32_t)__popc(M);
                    const uint32_t l = i < C ? ps.ol[ord] : start;
                    const uint32_t t = i < C ? i - ps.pre[l] : 0u;
                    const uint32_t c1 = ps.c1[l], c2 = ps.c2[l];
                    kk[u] = t < c1 ? ps.lo[0][l] + t : (t < c2 ? ps.lo[1][l] + (t - c1) : ps.lo[2][l] + (t - c2));
                    own[u] = l;
                    h[u] = __ldg(bp.hot + (i < C ? kk[u] : 0u));   // (index 0 always exists: the arrays are padded by one record)
                }
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const uint32_t i = i0 + 32u * u + lane;
                    const float4 a = ps.sa[own[u]];
                    const uint32_t oslot = __float_as_uint(h[u].w) & HOT_SLOT_MASK;
                    const float dx = a.x - h[u].x, dy = a.y - h[u].y;
                    const float d2 = __fmaf_rn(dx, dx, dy * dy);
                    const float mdk = (a.z + h[u].z) * 1.00005f;   // prefilter, see gather_single
                    const bool pass = i < C && oslot != (ps.sb[own[u]].w & HOT_SLOT_MASK) && !(d2 > mdk * mdk);
                    const uint32_t bl = __ballot_sync(FULL, pass);
                    const uint32_t w = qn + (uint32_t)__popc(bl & (le_mask >> 1));
                    if (pass && w < (uint32_t)COOP_Q) {
                        ps.key[w] = kk[u];
                        ps.owner[w] = (uint8_t)own[u];
                    }
                    qn += (uint32_t)__popc(bl);
                }
                if (qn > (uint32_t)COOP_Q) { ovf = true; break; }
            }
            if (!ovf) break;
            end = start + max(1u, (end - start) >> 1);   // more survivors than the queue holds: redo with half the lanes
        }
        const bool mine = coop && lane >= start && lane < end;
        __syncwarp();
        // first / last queue entry of every owner (the queue is in candidate order, so an owner's entries are contiguous)
        for (uint32_t i0 = 0; i0 < qn; i0 += 32u) {
            const uint32_t i = i0 + lane;
            const uint32_t o = i < qn ? (uint32_t)ps.owner[i] : 0xffu;
            const uint32_t prev = __shfl_up_sync(FULL, o, 1), next = __shfl_down_sync(FULL, o, 1);
            const uint32_t before = lane == 0u ? (i0 ? (uint32_t)ps.owner[i0 - 1u] : 0xfeu) : prev;
            const uint32_t after = lane == 31u ? (i + 1u < qn ? (uint32_t)ps.owner[i + 1u] : 0xfeu) : next;
            if (i < qn && o != before) ps.seg[o] = i;
            if (i < qn && o != after) ps.send[o] = i + 1u;
        }
        __syncwarp();
        // ---- 3. exact narrowphase, one (owner, candidate) pair per lane ---------------------------------------------------
        for (uint32_t i0 = 0; i0 < qn; i0 += 32u) {
            const uint32_t i = i0 + lane;
            if (i < qn) {
                const uint32_t o = (uint32_t)ps.owner[i];
                const float4 a = ps.sa[o];
                const uint4 bb = ps.sb[o];
                SelfCol so;
                so.x = so.qx = a.x; so.y = so.qy = a.y; so.r = a.z; so.m = a.w;
                so.memb = bb.x; so.filt = bb.y; so.body = bb.z; so.slot = bb.w & HOT_SLOT_MASK; so.sensor = (bb.w & HOT_SENSOR_BIT) != 0u; so.wbase = 0u;
                const Rec r = rec_of(__ldg(bp.hot + ps.key[i]), ccold);
                Contact ct;
                uint32_t key = POOL_NONE;
                if (narrowphase(so, r, ct)) {
                    note_pair(so, r, ct, out, rec, vel, stats);
                    if (ct.coincident) {
                        atomicOr(&ps.fb, 1u << o);
                    } else if (ct.push) {
                        key = pair_key<uint32_t>(so.slot, ct.other, false);
                        ps.cx[i] = ct.cx;
                        ps.cy[i] = ct.cy;
                        atomicAdd(&ps.nvalid[o], 1u);
                    }
                }
                ps.key[i] = key;
            }
        }
        __syncwarp();
        // ---- 4. rank of every contribution inside its owner's run -----------------------------------------------------------
        for (uint32_t i0 = 0; i0 < qn; i0 += 32u) {
            const uint32_t i = i0 + lane;
            if (i < qn) {
                const uint32_t key = ps.key[i];
                if (key != POOL_NONE) {
                    const uint32_t o = (uint32_t)ps.owner[i];
                    const uint32_t a = ps.seg[o], e = ps.send[o];
                    uint32_t rank = 0;
#pragma unroll 4
                    for (uint32_t j = a; j < e; ++j) rank += ps.key[j] < key ? 1u : 0u;
                    ps.perm[a + rank] = (uint16_t)i;
                }
            }
        }
        __syncwarp();
        // ---- 5. every owner adds its contributions in rank order ---------------------------------------------------------
        if (mine) {
            applied = true;
            if (!((ps.fb >> lane) & 1u)) {
                const uint32_t nv = ps.nvalid[lane], a = ps.seg[lane];
                for (uint32_t r = 0; r < nv; ++r) {
                    const uint32_t e = ps.perm[a + r];
                    px = fadd(px, ps.cx[e]);
                    py = fadd(py, ps.cy[e]);
                }
            }
        }
        start = end;
    }
    __syncwarp();
    if (applied && ((ps.fb >> lane) & 1u)) {   // coincident pair seen: exact serial path (pairs were already counted)
        SelfCol s2 = s;
        const float2 q = apply_contacts_rescan(g, bp, ccold, &s2, 1, px, py);
        px = q.x;
        py = q.y;
    }
    if (valid && !coop) {   // wrapped / tall cell range, or more candidates than the queue holds
        // really big neighbourhoods (the boundary shell of cfg2: hundreds of candidates) go to k_crowded unscanned - a whole warp
        // per body beats one lane
        if (defer_big && plain) big = true;
        else gather_single<true, uint32_t, 4>(g, bp, ccold, s, list, out, rec, vel, stats);
    }
    return applied;
}

// ------------------------------------------------------------------------------------------------
// update_objects for one body (physics.rs:326-358) + gravity (physics.rs:369-375) + circle constraints
// (physics.rs:377-395). (px, py) is the body position after contacts/joints; po / a / hv / gmod are the body's
// position_old, acceleration, velocity_request flag and gravity_mod, pre-loaded by the caller so that their memory
// latency overlaps the contact pass. Returns the pre-clamp position in (sx, sy): that is what the collider snapshot is
// built from (physics.rs:360-366 runs before apply_constraints).
// ------------------------------------------------------------------------------------------------
// acceleration / pending velocity_request of a body as update_objects will see them (SubstepParams::acc_zero: known to be none)
__device__ __forceinline__ float2 load_acc(const SubstepParams& P, const BodyArrays& B, uint32_t b) {
    return P.acc_zero ? make_float2(0.f, 0.f) : B.acc[b];
}
__device__ __forceinline__ bool load_hv(const SubstepParams& P, const BodyArrays& B, uint32_t b) { return P.acc_zero ? false : B.has_vreq[b] != 0; }

__device__ __forceinline__ void integrate_body(const SubstepParams& P, const Constraints& K, const BodyArrays& B, uint32_t b,
                                               uint32_t flags, float gmod, float px, float py, float2 po, float2 a, bool hv,
                                               float& sx, float& sy, float& rot_out, DeviceStats* stats) {
    float rot = 0.0f;
    if (flags & BF_ROT) rot = B.rot[b];
    if (flags & BF_STATIC) {                                     // physics.rs:327-332
        B.pos_old[b] = make_float2(px, py);
        if (!P.acc_zero) B.acc[b] = make_float2(0.f, 0.f);
        if (P.write_vel) B.vel[b] = make_float2(0.f, 0.f);
    } else {
        if (hv) {                                                // physics.rs:334-336
            const float2 v = B.vreq[b];
            po.x = fsub(px, fmul(v.x, P.dt));
            po.y = fsub(py, fmul(v.y, P.dt));
            B.has_vreq[b] = 0;
        }
        const float ratio = (flags & BF_FIRST_DYN) ? P.ratio_first : P.ratio_rest;   // physics.rs:338-339
        const float dx = fmul(fsub(px, po.x), ratio), dy = fmul(fsub(py, po.y), ratio);
        if (!(flags & BF_SPRINGS)) {                             // gravity; spring bodies got it in k_springs
            a.x = fadd(a.x, fmul(P.gx, gmod));
            a.y = fadd(a.y, fmul(P.gy, gmod));
        }
        B.pos_old[b] = make_float2(px, py);                      // physics.rs:343
        px = fadd(px, fadd(dx, fmul(fmul(a.x, P.dt), P.dt)));    // physics.rs:344
        py = fadd(py, fadd(dy, fmul(fmul(a.y, P.dt), P.dt)));
        if (flags & BF_ROT) {                                    // physics.rs:346-355
            float w = B.angvel[b];
            w = fadd(w, fmul(fdiv(B.torque[b], B.inertia[b]), P.dt));
            rot = fadd(rot, fmul(w, P.dt));
            B.angvel[b] = w;
            B.rot[b] = rot;
            B.torque[b] = 0.0f;
        }
        if (!P.acc_zero) B.acc[b] = make_float2(0.f, 0.f);
        // calculated_velocity (physics.rs:357) is only observable through springs, events and host reads, so it is
        // materialised on the substeps where one of those can see it (host sets write_vel)
        if (P.write_vel) B.vel[b] = make_float2(fdiv(dx, P.dt), fdiv(dy, P.dt));
    }
    sx = px;
    sy = py;
    rot_out = rot;
    for (int i = 0; i < K.n; ++i) {                              // physics.rs:377-395
        const float4 k = __ldg(K.c + i);                         // (x, y, radius, -)
        const float tx = fsub(px, k.x), ty = fsub(py, k.y);
        const float d2 = fadd(fmul(tx, tx), fmul(ty, ty));
        // sqrt(d2) > R needs d2 > R*R*(1 - 1e-4) at the very least: skip the sqrt for bodies well inside the circle
        if (k.z < 0.f || d2 > k.z * k.z * 0.9999f) {
            const float d = __fsqrt_rn(d2);
            if (d > k.z) {
                px = fadd(k.x, fmul(fdiv(tx, d), k.z));
                py = fadd(k.y, fmul(fdiv(ty, d), k.z));
            }
        }
    }
    if (px != px || py != py) atomicOr(&stats->nan_flag, 1u);
    B.pos[b] = make_float2(px, py);
}

// ---- list pipeline: displacement tracking -------------------------------------------------------------------------
// Every publisher (a kernel that writes new collider snapshots) measures how far each snapshot has moved from where it was
// when the neighbour lists were built, relative to a common displacement c (bodies falling together have not moved RELATIVE
// to each other): m = |snapshot - ref - c|. A pair that is NOT in a list was further apart than r_a + r_b + skin at build time,
// and |(d_a - c) - (d_b - c)| <= m_a + m_b, so while max m <= 0.45 * skin no such pair can touch: the lists remain supersets
// of the contact set. The slack term (a few ulps of the coordinates) covers the rounding of this estimate and of the exact
// narrowphase. c is only an estimate (sampled mean, extrapolated by k_nl_decide): it decides WHEN lists are rebuilt, never
// WHAT the contacts are.
struct NlAcc { float m, dx, dy; uint32_t n; };
__device__ __forceinline__ void nl_track(float ax, float ay, float refx, float refy, float cx, float cy, NlAcc& na) {
    const float ddx = ax - refx, ddy = ay - refy;
    const float ex = ddx - cx, ey = ddy - cy;
    float m = sqrtf(ex * ex + ey * ey) * 1.000001f + 4e-6f * (fabsf(ax) + fabsf(ay));
    if (!(m == m)) m = 3.4e38f;   // NaN position: forces a rebuild every substep (lists then always date from this very snapshot)
    na.m = fmaxf(na.m, m);
    if (ddx == ddx && ddy == ddy && fabsf(ddx) < 1e30f && fabsf(ddy) < 1e30f) { na.dx += ddx; na.dy += ddy; na.n++; }
}
// CTA-wide commit of the tracking (and of a pair count): one atomicMax per CTA and only if it raises the maximum; every 64th
// CTA contributes its displacement sum (a sample is enough for an estimate). Must be reached by every thread of the CTA.
__device__ __forceinline__ void nl_commit(NlCtl* ctl, const NlAcc& na, unsigned long long* pair_counter, unsigned int n_pairs) {
    __shared__ uint32_t s_m[32], s_n[32], s_p[32];
    __shared__ float s_x[32], s_y[32];
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
    const bool sample = (blockIdx.x & 63u) == 0u;
    const uint32_t mb = __reduce_max_sync(FULL, __float_as_uint(na.m));
    const uint32_t np = __reduce_add_sync(FULL, n_pairs);
    float sx = na.dx, sy = na.dy;
    uint32_t sn = na.n;
    if (sample) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { sx += __shfl_xor_sync(FULL, sx, d); sy += __shfl_xor_sync(FULL, sy, d); }
        sn = __reduce_add_sync(FULL, sn);
    }
    if (lane == 0) { s_m[warp] = mb; s_p[warp] = np; s_x[warp] = sx; s_y[warp] = sy; s_n[warp] = sn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t M = 0, N = 0, Pn = 0;
        float X = 0.f, Y = 0.f;
        for (uint32_t w = 0; w < nw; ++w) { M = max(M, s_m[w]); Pn += s_p[w]; X += s_x[w]; Y += s_y[w]; N += s_n[w]; }
        if (Pn) atomicAdd(pair_counter, (unsigned long long)Pn);
        if (ctl != nullptr) {
            if (M > *reinterpret_cast<volatile unsigned int*>(&ctl->max_m)) atomicMax(&ctl->max_m, M);
            if (sample && N) { atomicAdd(&ctl->sum_x, X); atomicAdd(&ctl->sum_y, Y); atomicAdd(&ctl->n_sum, N); }
        }
    }
}

// Collider snapshot (physics.rs:360-366): abs.translation = M(rot) * offset.translation + pos, with
// glam's Mat2::from_angle columns (cos, sin), (-sin, cos) and M*v = x_axis*v.x + y_axis*v.y.
__device__ __forceinline__ float2 snapshot_of(const ColliderArrays& Cc, uint32_t c, uint32_t cflags, float sx, float sy, float rot) {
    float2 off = make_float2(0.f, 0.f);
    if (cflags & CF_OFFSET) off = Cc.coff[c];
    float sn = 0.0f, cs = 1.0f;
    if (rot != 0.0f) sincosf(rot, &sn, &cs);
    return make_float2(fadd(fadd(fmul(cs, off.x), fmul(-sn, off.y)), sx), fadd(fadd(fmul(sn, off.x), fmul(cs, off.y)), sy));
}

// Bins one collider: rank within its cell (atomic on the cell counter) + warp-aggregated add to the counter of the
// scan tile that owns the cell (lanes of a warp mostly share a tile, so this is ~1 extra atomic per warp).
__device__ __forceinline__ uint32_t bin_collider(uint32_t* tab_next, uint32_t* tile_next, uint32_t cell) {
    const uint32_t rank = atomicAdd(tab_next + cell, 1u);
    const uint32_t tile = cell >> SCAN_TILE_SHIFT;
    const unsigned int peers = __match_any_sync(__activemask(), tile);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(tile_next + tile, (uint32_t)__popc(peers));
    return rank;
}

// Publishes the new snapshot of one collider. Grid pipeline: bins it into the table under construction. List pipeline: writes
// the slot-indexed record for the next substep's contact pass and tracks its displacement (see nl_track).
__device__ __forceinline__ float2 publish_collider(const GridDesc& g, const ColliderArrays& Cc, const Broadphase& bp, uint32_t c, uint32_t cflags,
                                                   uint32_t wbase, float sx, float sy, float rot, NlAcc& na) {
    const float2 a = snapshot_of(Cc, c, cflags, sx, sy, rot);
    Cc.cabs[c] = a;
    if (bp.nl.snap_next != nullptr) {
        const float4 me = __ldg(bp.nl.snap_cur + c);
        const uint4 hd = bp.nl.hdr[c];
        const float4 rec_new = make_float4(a.x, a.y, me.z, me.w);
        bp.nl.snap_next[c] = rec_new;
        if (hd.w & (NLF_PUSH_L | NLF_PUSH_R)) {   // strips: a neighbour rank keeps this collider as a ghost
            if (hd.w & NLF_PUSH_L) bp.nl.peer_next[0][c] = rec_new;
            if (hd.w & NLF_PUSH_R) bp.nl.peer_next[1][c] = rec_new;
            __threadfence_system();
        }
        nl_track(a.x, a.y, __uint_as_float(hd.x), __uint_as_float(hd.y), bp.nl.ctl->cx, bp.nl.ctl->cy, na);
    } else {
        const uint32_t cell = wbase + cell_index(g, bin_coord(a.x, g.inv_cell), bin_coord(a.y, g.inv_cell));
        Cc.ccell[c] = make_uint2(cell, bin_collider(bp.tab_next, bp.tile_next, cell));
    }
    return a;
}
// list pipeline: kernels that walk the cell grid (k_multi, k_crowded) use the tables of the last rebuild
__device__ __forceinline__ Broadphase resolve_grid(Broadphase bp) {
    if (bp.nl.snap_next != nullptr) {
        bp.tab = (bp.nl.ctl->parity & 1u) ? bp.nl.tab[1] : bp.nl.tab[0];
        bp.hot = bp.nl.hot;
        bp.snap = bp.nl.snap_cur;
    }
    return bp;
}

// ---- strip decomposition: message packing (see the strip section at the end of this file) --------------------------
__device__ __forceinline__ void strip_append_ghost(void* msg, const StripDesc& S, float4 hot) {
    StripHeader* h = reinterpret_cast<StripHeader*>(msg);
    const uint32_t i = atomicAdd(&h->n_ghost, 1u);
    if (i < S.gcap) strip_ghosts(msg)[i] = hot;
    else h->overflow = 1u;
}

// One owned collider with its NEW snapshot a: ghost record for whichever neighbour can reach it, full body state if it left
// the strip. A neighbour-owned partner at x' >= x_hi needs me iff x' - x < r + r' <= reach (directed rounding keeps the
// selection conservative).
__device__ __forceinline__ void strip_pack_one(const BodyArrays& B, const ColliderArrays& Cc, const StripDesc& S, uint32_t c, uint32_t cflags,
                                               float2 a, float r, void* send_l, void* send_r) {
    const float reach = __fadd_ru(r, S.rmax);
    const float4 hot = make_float4(a.x, a.y, r, __uint_as_float(hot_word(c, cflags)));
    if (S.has_right && a.x >= __fsub_rd(S.x_hi, reach)) strip_append_ghost(send_r, S, hot);
    if (S.has_left && a.x < __fadd_ru(S.x_lo, reach)) strip_append_ghost(send_l, S, hot);
    void* dst = nullptr;
    if (S.has_right && a.x >= S.x_hi) dst = send_r;
    else if (S.has_left && a.x < S.x_lo) dst = send_l;
    if (dst != nullptr) {
        StripHeader* h = reinterpret_cast<StripHeader*>(dst);
        const uint32_t i = atomicAdd(&h->n_mig, 1u);
        if (i < S.mcap) {
            const uint32_t b = Cc.cparent[c];
            MigRec m;
            m.slot = b; m.col = c;
            m.pos = B.pos[b]; m.pos_old = B.pos_old[b]; m.acc = B.acc[b]; m.vel = B.vel[b]; m.vreq = B.vreq[b]; m.cabs = a;
            m.rot = B.rot[b]; m.angvel = B.angvel[b]; m.torque = B.torque[b]; m.has_vreq = B.has_vreq[b];
            m.pad[0] = m.pad[1] = 0u;
            strip_migs(dst, S.gcap)[i] = m;
        } else {
            h->overflow = 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K-main: one thread per body slot. Contacts (gather, ordered) [+ verlet + snapshot + clamp + binning when FUSED].
// Handles bodies with zero or one collider; multi-collider bodies go to k_multi.
// Memory-latency chain (the kernel is latency-, not bandwidth-bound): round 1 = every per-body array plus the collider
// arrays at the SPECULATED slot c == b (insert_rbd + insert_collider_with_parent in lockstep gives identical slots);
// round 2 = six cell-table entries; round 3 = hot record halves; round 4 = cold halves of prefilter survivors;
// round 5 = the binning atomic.
// ------------------------------------------------------------------------------------------------
template <bool FUSED, bool ORDERED, int BATCH, int MINB, bool POOLED, int THREADS = 256>
__global__ void __launch_bounds__(THREADS, MINB) k_main(SubstepParams P, GridDesc g, Constraints K, BodyArrays B, ColliderArrays Cc,
                                              Broadphase bp, Recording rec, DeviceStats* stats, StripView sv) {
    __shared__ CoopSmem pool[POOLED ? THREADS / 32 : 1];
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    GatherOut out;
    out.fx = out.fy = 0.f;
    out.n_pairs = out.n_coinc = 0;
    unsigned int n_over = 0;
    bool inb;
    if (sv.olist != nullptr) {  // strip mode: threads enumerate the compact list of bodies this rank owns
        inb = b < __ldg(sv.ocount);
        b = inb ? sv.olist[b] : NO_SLOT;
        inb = b != NO_SLOT;
    } else {
        inb = b < P.n_bodies;
    }
    const uint32_t bl = inb ? b : 0u;
    // round 1: everything that only needs b (tail threads read slot 0 and discard)
    const uint2 info = B.binfo[bl];
    const float2 mg = B.bmg[bl];
    float2 p = B.pos[bl];
    const float2 po = B.pos_old[bl];
    const float2 acc0 = load_acc(P, B, bl);
    const bool hv = load_hv(P, B, bl);
    const uint32_t wbase = g.n_worlds > 1u ? B.bworld[bl] * g.ncells : 0u;
    uint32_t cs = 0;
    uint4 cc = make_uint4(0u, 0u, 0u, 0u);
    float2 ab = make_float2(0.f, 0.f);
    if (P.n_colliders) {
        cs = min(bl, P.n_colliders - 1u);
        cc = Cc.cconst[cs];
        ab = Cc.cabs[cs];
    }
    const uint32_t flags = info.x;
    const int32_t col = (int32_t)info.y;
    if (!POOLED) {   // per-lane contact resolution (the variant tuned for sparse contacts; kept textually as it was)
        if (inb && (flags & BF_ALIVE) && col >= BODY_NO_COLLIDER) {
            bool active_col = false, deferred = false;
            if (col >= 0) {
                const uint32_t c = (uint32_t)col;
                if (c != cs) {  // speculation missed: fetch the real collider
                    cc = Cc.cconst[c];
                    ab = Cc.cabs[c];
                }
                active_col = (cc.y & CF_ACTIVE) != 0u;
                if (active_col && P.collisions_enabled) {
                    SelfCol s;
                    s.x = s.qx = ab.x; s.y = s.qy = ab.y; s.r = __uint_as_float(cc.x); s.m = mg.x;
                    s.memb = cc.z; s.filt = cc.w; s.body = b; s.slot = c; s.wbase = wbase; s.sensor = (cc.y & CF_SENSOR) != 0u;
                    ContactList<uint32_t> list;
                    list.clear();
                    gather_single<ORDERED, uint32_t, BATCH>(g, bp, Cc.ccold, s, list, out, rec, B.vel, stats);
                    if (ORDERED) {
                        if (!list.overflow) {
                            for (int i = 0; i < list.n; ++i) { p.x = fadd(p.x, list.cx[i]); p.y = fadd(p.y, list.cy[i]); }
                        } else {
                            n_over = 1;
                            if (P.crowded) {  // a whole warp of k_crowded redoes this body, including the fused tail below
                                P.over_list[atomicAdd(&stats->over_count[P.over_parity], 1u)] = b;
                                deferred = true;
                            } else {
                                SelfCol s2 = s;  // stack copy only on this rare path
                                p = apply_contacts_rescan(g, bp, Cc.ccold, &s2, 1, p.x, p.y);
                            }
                        }
                    } else {
                        p.x = fadd(p.x, out.fx);
                        p.y = fadd(p.y, out.fy);
                    }
                }
            }
            if (deferred) {
                // nothing: every array of this body is left untouched for k_crowded
            } else if (FUSED && !(flags & BF_JOINTED)) {   // jointed bodies are advanced by k_joints_fused after the joint projection
                float sx, sy, rot;
                integrate_body(P, K, B, b, flags, mg.y, p.x, p.y, po, acc0, hv, sx, sy, rot, stats);
                if (active_col) {
                    NlAcc na_unused{0.f, 0.f, 0.f, 0u};
                    const float2 a = publish_collider(g, Cc, bp, (uint32_t)col, cc.y, wbase, sx, sy, rot, na_unused);
                    if (sv.olist != nullptr) strip_pack_one(B, Cc, sv.S, (uint32_t)col, cc.y, a, __uint_as_float(cc.x), sv.send_l, sv.send_r);
                }
            } else {
                B.pos[b] = p;
            }
        }
    } else {
        const bool do_body = inb && (flags & BF_ALIVE) && col >= BODY_NO_COLLIDER;
        bool active_col = false, deferred = false, do_gather = false;
        SelfCol s;
        s.x = s.y = s.r = s.m = s.qx = s.qy = 0.f;
        s.memb = s.filt = 0u; s.body = b; s.slot = 0u; s.wbase = wbase; s.sensor = false;
        if (do_body && col >= 0) {
            const uint32_t c = (uint32_t)col;
            if (c != cs) {  // speculation missed: fetch the real collider
                cc = Cc.cconst[c];
                ab = Cc.cabs[c];
            }
            active_col = (cc.y & CF_ACTIVE) != 0u;
            if (active_col && P.collisions_enabled) {
                s.x = s.qx = ab.x; s.y = s.qy = ab.y; s.r = __uint_as_float(cc.x); s.m = mg.x;
                s.memb = cc.z; s.filt = cc.w; s.body = b; s.slot = c; s.wbase = wbase; s.sensor = (cc.y & CF_SENSOR) != 0u;
                do_gather = true;
            }
        }
        if (ORDERED) {
            ContactList<uint32_t> list;
            list.clear();
            bool applied = false, big = false;
            if (POOLED) {   // warp-collective: every lane calls
                applied = gather_coop<BATCH>(g, bp, Cc.ccold, do_gather, s, list, out, rec, B.vel, stats, pool[threadIdx.x >> 5], P.crowded != 0u, big, p.x, p.y);
            } else if (do_gather) {
                gather_single<true, uint32_t, BATCH>(g, bp, Cc.ccold, s, list, out, rec, B.vel, stats);
            }
            if (big) {   // k_crowded does the whole body, pair counting included
                n_over = 1;
                P.over_list[atomicAdd(&stats->over_count[P.over_parity], 1u)] = OVER_COUNT_BIT | b;
                deferred = true;
            } else if (do_gather && !applied) {
                if (!list.overflow) {
                    for (int i = 0; i < list.n; ++i) { p.x = fadd(p.x, list.cx[i]); p.y = fadd(p.y, list.cy[i]); }
                } else {
                    n_over = 1;
                    if (P.crowded) {  // a whole warp of k_crowded redoes this body, including the fused tail below
                        P.over_list[atomicAdd(&stats->over_count[P.over_parity], 1u)] = b;
                        deferred = true;
                    } else {
                        SelfCol s2 = s;  // stack copy only on this rare path
                        p = apply_contacts_rescan(g, bp, Cc.ccold, &s2, 1, p.x, p.y);
                    }
                }
            }
        } else if (do_gather) {
            ContactList<uint32_t> list;
            list.clear();
            gather_single<false, uint32_t, BATCH>(g, bp, Cc.ccold, s, list, out, rec, B.vel, stats);
            p.x = fadd(p.x, out.fx);
            p.y = fadd(p.y, out.fy);
        }
        if (do_body) {
            if (deferred) {
                // nothing: every array of this body is left untouched for k_crowded
            } else if (FUSED && !(flags & BF_JOINTED)) {   // jointed bodies are advanced by k_joints_fused after the joint projection
                float sx, sy, rot;
                integrate_body(P, K, B, b, flags, mg.y, p.x, p.y, po, acc0, hv, sx, sy, rot, stats);
                if (active_col) {
                    NlAcc na_unused{0.f, 0.f, 0.f, 0u};
                    const float2 a = publish_collider(g, Cc, bp, (uint32_t)col, cc.y, wbase, sx, sy, rot, na_unused);
                    if (sv.olist != nullptr) strip_pack_one(B, Cc, sv.S, (uint32_t)col, cc.y, a, __uint_as_float(cc.x), sv.send_l, sv.send_r);
                }
            } else {
                B.pos[b] = p;
            }
        }
    }
    warp_add_u64(&stats->collisions, out.n_pairs);
    if (__any_sync(0xffffffffu, out.n_coinc | n_over)) {
        warp_add_u64(&stats->coincident, out.n_coinc);
        unsigned int o = __reduce_add_sync(0xffffffffu, n_over);
        if ((threadIdx.x & 31) == 0 && o) atomicAdd(&stats->list_overflow, o);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA bulk-copy helpers (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS). Round 2 measured a tile kernel that staged its
// candidate windows this way (profiles/r2_ktile_tma_verdict.md): bit-exact, but no faster than plain loads for this gather-bound
// access pattern, so that kernel was removed; the helpers stay for contiguous per-CTA state tiles.
// ------------------------------------------------------------------------------------------------
#ifndef BLOBS_EMU
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
#endif

// ------------------------------------------------------------------------------------------------
// K-multi: bodies with more than one distinct collider. One thread per such body; the contributions of all its
// colliders are merged into one ordered list (order = (later slot, earlier slot) over the union).
// ------------------------------------------------------------------------------------------------
constexpr int MULTI_MAX_INLINE = 8;  // colliders staged for the rescan path

__device__ __forceinline__ bool load_self(const BodyArrays& B, const ColliderArrays& Cc, const Broadphase& bp, uint32_t b, uint32_t c, float m,
                                          uint32_t wbase, SelfCol& s) {
    const uint4 cc = Cc.cconst[c];
    if (!(cc.y & CF_ACTIVE)) return false;
    const float2 a = Cc.cabs[c];
    s.x = s.qx = a.x; s.y = s.qy = a.y; s.r = __uint_as_float(cc.x); s.m = m;
    if (bp.snap != nullptr) {   // list pipeline: the grid was built around the reference position
        const uint4 hd = bp.nl.hdr[c];
        s.qx = __uint_as_float(hd.x);
        s.qy = __uint_as_float(hd.y);
    }
    s.memb = cc.z; s.filt = cc.w; s.body = b; s.slot = c; s.wbase = wbase; s.sensor = (cc.y & CF_SENSOR) != 0u;
    return true;
}

// Serial resolution of a multi-collider body with more than LIST_CAP contributions: order-preserving windowed rescan over
// all colliders of the body; bodies with more than MULTI_MAX_INLINE colliders fall back to the unordered sum (counted in
// list_overflow).
__device__ __forceinline__ float2 multi_overflow_serial(const GridDesc& g, const Broadphase& bp, const BodyArrays& B, const ColliderArrays& Cc,
                                                        uint32_t b, const uint32_t* __restrict__ mb_cols, uint32_t c0, uint32_t c1, float m,
                                                        uint32_t wbase, float2 p) {
    SelfCol cols[MULTI_MAX_INLINE];
    int nc = 0;
    bool fits = true;
    for (uint32_t k = c0; k < c1; ++k) {
        SelfCol s;
        if (!load_self(B, Cc, bp, b, mb_cols[k], m, wbase, s)) continue;
        if (nc == MULTI_MAX_INLINE) { fits = false; break; }
        cols[nc++] = s;
    }
    if (fits) return apply_contacts_rescan(g, bp, Cc.ccold, cols, nc, p.x, p.y);
    float fx = 0.f, fy = 0.f;
    for (uint32_t k = c0; k < c1; ++k) {
        SelfCol s;
        if (!load_self(B, Cc, bp, b, mb_cols[k], m, wbase, s)) continue;
        for_each_candidate(g, bp, Cc.ccold, s.wbase, s.qx, s.qy, s.r, [&](const Rec& o) {
            Contact c;
            if (!narrowphase(s, o, c)) return;
            if (c.coincident) fx = fadd(fx, c.i_am_a ? 0.01f : -0.01f);
            if (c.push) { fx = fadd(fx, c.cx); fy = fadd(fy, c.cy); }
        });
    }
    return make_float2(fadd(p.x, fx), fadd(p.y, fy));
}

template <bool FUSED, bool ORDERED>
__global__ void __launch_bounds__(128) k_multi(SubstepParams P, GridDesc g, Constraints K, BodyArrays B, ColliderArrays Cc,
                                               Broadphase bp_in, Recording rec, DeviceStats* stats, const uint32_t* __restrict__ mb_body,
                                               const uint32_t* __restrict__ mb_off, const uint32_t* __restrict__ mb_cols,
                                               uint32_t n_multi) {
    const Broadphase bp = resolve_grid(bp_in);
    NlAcc na{0.f, 0.f, 0.f, 0u};
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    GatherOut out;
    out.fx = out.fy = 0.f;
    out.n_pairs = out.n_coinc = 0;
    unsigned int n_over = 0;
    if (i < n_multi) {
        const uint32_t b = mb_body[i];
        const uint32_t flags = B.binfo[b].x;
        const float2 mg = B.bmg[b];
        const uint32_t c0 = mb_off[i], c1 = mb_off[i + 1];
        float2 p = B.pos[b];
        const float2 po = B.pos_old[b];
        const float2 acc0 = load_acc(P, B, b);
        const bool hv = load_hv(P, B, b);
        const uint32_t wbase = g.n_worlds > 1u ? B.bworld[b] * g.ncells : 0u;
        bool deferred = false;
        if (P.collisions_enabled) {
            ContactList<unsigned long long> list;
            list.clear();
            const float m = mg.x;
            for (uint32_t k = c0; k < c1; ++k) {
                SelfCol s;
                if (!load_self(B, Cc, bp, b, mb_cols[k], m, wbase, s)) continue;
                gather_generic<ORDERED, unsigned long long>(g, bp, Cc.ccold, s, list, out, rec, B.vel, stats);
            }
            if (ORDERED) {
                if (!list.overflow) {
                    for (int j = 0; j < list.n; ++j) { p.x = fadd(p.x, list.cx[j]); p.y = fadd(p.y, list.cy[j]); }
                } else {
                    n_over = 1;
                    if (P.crowded) {
                        P.over_list[atomicAdd(&stats->over_count[P.over_parity], 1u)] = OVER_MULTI_BIT | i;
                        deferred = true;
                    } else {
                        p = multi_overflow_serial(g, bp, B, Cc, b, mb_cols, c0, c1, m, wbase, p);
                    }
                }
            } else {
                p.x = fadd(p.x, out.fx);
                p.y = fadd(p.y, out.fy);
            }
        }
        if (deferred) {
            // k_crowded finishes this body
        } else if (FUSED && !(flags & BF_JOINTED)) {
            float sx, sy, rot;
            integrate_body(P, K, B, b, flags, mg.y, p.x, p.y, po, acc0, hv, sx, sy, rot, stats);
            for (uint32_t k = c0; k < c1; ++k) {
                const uint32_t c = mb_cols[k];
                const uint32_t cf = Cc.cconst[c].y;
                if (cf & CF_ACTIVE) publish_collider(g, Cc, bp, c, cf, wbase, sx, sy, rot, na);
            }
        } else {
            B.pos[b] = p;
        }
    }
    if (bp.nl.snap_next != nullptr) nl_commit(bp.nl.ctl, na, &stats->collisions, 0u);
    warp_add_u64(&stats->collisions, out.n_pairs);
    if (__any_sync(0xffffffffu, out.n_coinc | n_over)) {
        warp_add_u64(&stats->coincident, out.n_coinc);
        unsigned int o = __reduce_add_sync(0xffffffffu, n_over);
        if ((threadIdx.x & 31) == 0 && o) atomicAdd(&stats->list_overflow, o);
    }
}

// ------------------------------------------------------------------------------------------------
// K-crowded: bodies whose ordered contact list overflowed in k_main / k_multi (more than LIST_CAP contributions), one WARP
// per body. Such bodies appear where the circle constraint projects many bodies onto its boundary: a few hundred threads
// with hundreds of contacts each would otherwise serialise the tail of k_main. The lanes walk the neighbourhood together
// (lane-strided over each row span), drop their contributions into a shared-memory buffer, the warp sorts it by pair key
// (bitonic) and every lane replays the sum in reference order; lane 0 then runs the same tail as k_main / k_multi
// (verlet + snapshot + clamp + binning [+ strip packing]). Pair counting / recording already happened in the first pass.
// More than CROWD_CAP contributions: lane 0 falls back to the serial windowed rescan. The pooled k_main also sends bodies with
// more than 64 candidates straight here (OVER_COUNT_BIT: their pairs are counted / recorded by this kernel).
// over_count is double-buffered by substep parity: this launch consumes [parity] and clears [parity ^ 1] for the next substep.
// ------------------------------------------------------------------------------------------------
constexpr int CROWD_CAP = 512;
constexpr int CROWD_WARPS = 2;

template <class F>
__device__ __forceinline__ void warp_for_each_candidate(const GridDesc& g, const Broadphase& bp, const uint4* __restrict__ ccold, uint32_t wbase,
                                                        float x, float y, float r, uint32_t lane, F&& f) {
    const CellRange R = cell_range(g, x, y, r);
    const uint32_t n1 = min(R.nx, g.W - R.c0);
    for (uint32_t j = 0; j < R.ny; ++j) {
        uint32_t row = R.r0 + j;
        if (row >= g.H) row -= g.H;
        const uint32_t base = wbase + row * g.W;
        uint32_t lo = __ldg(bp.tab + base + R.c0), hi = __ldg(bp.tab + base + R.c0 + n1);
        for (uint32_t k = lo + lane; k < hi; k += 32u) f(load_rec(bp, ccold, k));
        if (n1 < R.nx) {
            lo = __ldg(bp.tab + base);
            hi = __ldg(bp.tab + base + (R.nx - n1));
            for (uint32_t k = lo + lane; k < hi; k += 32u) f(load_rec(bp, ccold, k));
        }
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(32 * CROWD_WARPS) k_crowded(SubstepParams P, GridDesc g, Constraints K, BodyArrays B, ColliderArrays Cc,
                                                              Broadphase bp_in, Recording rec, DeviceStats* stats, StripView sv,
                                                              const uint32_t* __restrict__ mb_body, const uint32_t* __restrict__ mb_off,
                                                              const uint32_t* __restrict__ mb_cols) {
    const Broadphase bp = resolve_grid(bp_in);
    NlAcc na{0.f, 0.f, 0.f, 0u};
    GatherOut out;
    out.fx = out.fy = 0.f;
    out.n_pairs = out.n_coinc = 0;
    __shared__ unsigned long long skey[CROWD_WARPS][CROWD_CAP];
    __shared__ float scx[CROWD_WARPS][CROWD_CAP];
    __shared__ float scy[CROWD_WARPS][CROWD_CAP];
    __shared__ uint32_t scount[CROWD_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned long long* const key = skey[warp];
    float* const cx = scx[warp];
    float* const cy = scy[warp];
    uint32_t* const cnt = &scount[warp];
    if (blockIdx.x == 0 && threadIdx.x == 0) stats->over_count[P.over_parity ^ 1u] = 0u;
    const uint32_t n_over = min(*reinterpret_cast<volatile unsigned int*>(&stats->over_count[P.over_parity]), P.n_bodies);
    for (uint32_t w = blockIdx.x * CROWD_WARPS + warp; w < n_over; w += gridDim.x * CROWD_WARPS) {
        const uint32_t entry = P.over_list[w];
        const bool multi = (entry & OVER_MULTI_BIT) != 0u;
        const bool count = (entry & OVER_COUNT_BIT) != 0u;   // k_main skipped this body altogether: count / record its pairs here
        const uint32_t idx = entry & ~(OVER_MULTI_BIT | OVER_COUNT_BIT);
        const uint32_t b = multi ? mb_body[idx] : idx;
        const uint2 info = B.binfo[b];
        const uint32_t flags = info.x;
        const float2 mg = B.bmg[b];
        float2 p = B.pos[b];
        const uint32_t wbase = g.n_worlds > 1u ? B.bworld[b] * g.ncells : 0u;
        const uint32_t c0 = multi ? mb_off[idx] : 0u, c1 = multi ? mb_off[idx + 1] : 1u;
        if (lane == 0) *cnt = 0u;
        __syncwarp();
        for (uint32_t k = c0; k < c1; ++k) {
            SelfCol s;
            if (!load_self(B, Cc, bp, b, multi ? mb_cols[k] : info.y, mg.x, wbase, s)) continue;
            warp_for_each_candidate(g, bp, Cc.ccold, s.wbase, s.qx, s.qy, s.r, lane, [&](const Rec& o) {
                Contact c;
                if (!narrowphase(s, o, c)) return;
                if (count) note_pair(s, o, c, out, rec, B.vel, stats);
                if (c.coincident) {
                    const uint32_t i = atomicAdd(cnt, 1u);
                    if (i < (uint32_t)CROWD_CAP) { key[i] = pair_key<unsigned long long>(s.slot, c.other, true); cx[i] = c.i_am_a ? 0.01f : -0.01f; cy[i] = 0.f; }
                }
                if (c.push) {
                    const uint32_t i = atomicAdd(cnt, 1u);
                    if (i < (uint32_t)CROWD_CAP) { key[i] = pair_key<unsigned long long>(s.slot, c.other, false); cx[i] = c.cx; cy[i] = c.cy; }
                }
            });
        }
        __syncwarp();
        const uint32_t n = *cnt;
        if (n <= (uint32_t)CROWD_CAP) {
            uint32_t m = 2u;
            while (m < n) m <<= 1;
            for (uint32_t i = n + lane; i < m; i += 32u) key[i] = ~0ull;  // pads sort to the end (real keys are < 2^63)
            __syncwarp();
            for (uint32_t kk = 2u; kk <= m; kk <<= 1) {
                for (uint32_t j = kk >> 1; j > 0u; j >>= 1) {
                    for (uint32_t i = lane; i < m; i += 32u) {
                        const uint32_t l = i ^ j;
                        if (l > i) {
                            const unsigned long long ka = key[i], kb = key[l];
                            const bool up = (i & kk) == 0u;
                            if ((ka > kb) == up) {
                                key[i] = kb; key[l] = ka;
                                const float xa = cx[i], ya = cy[i];
                                cx[i] = cx[l]; cy[i] = cy[l];
                                cx[l] = xa; cy[l] = ya;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            for (uint32_t i = 0; i < n; ++i) { p.x = fadd(p.x, cx[i]); p.y = fadd(p.y, cy[i]); }  // every lane, same order
        } else if (lane == 0) {
            if (multi) {
                p = multi_overflow_serial(g, bp, B, Cc, b, mb_cols, c0, c1, mg.x, wbase, p);
            } else {
                SelfCol s;
                if (load_self(B, Cc, bp, b, info.y, mg.x, wbase, s)) p = apply_contacts_rescan(g, bp, Cc.ccold, &s, 1, p.x, p.y);
            }
        }
        if (lane == 0) {
            if (FUSED && !(flags & BF_JOINTED)) {
                const float2 po = B.pos_old[b];
                const float2 acc0 = load_acc(P, B, b);
                const bool hv = load_hv(P, B, b);
                float sx, sy, rot;
                integrate_body(P, K, B, b, flags, mg.y, p.x, p.y, po, acc0, hv, sx, sy, rot, stats);
                for (uint32_t k = c0; k < c1; ++k) {
                    const uint32_t c = multi ? mb_cols[k] : info.y;
                    const uint4 cc = Cc.cconst[c];
                    if (!(cc.y & CF_ACTIVE)) continue;
                    const float2 a = publish_collider(g, Cc, bp, c, cc.y, wbase, sx, sy, rot, na);
                    if (!multi && sv.olist != nullptr) strip_pack_one(B, Cc, sv.S, c, cc.y, a, __uint_as_float(cc.x), sv.send_l, sv.send_r);
                }
            } else {
                B.pos[b] = p;
            }
        }
        __syncwarp();
    }
    if (bp.nl.snap_next != nullptr) nl_commit(bp.nl.ctl, na, &stats->collisions, 0u);
    warp_add_u64(&stats->collisions, out.n_pairs);
    if (__any_sync(0xffffffffu, out.n_coinc)) warp_add_u64(&stats->coincident, out.n_coinc);
}

// ------------------------------------------------------------------------------------------------
// K-integrate (split pipeline, used when joints or event recording forbid fusion): update_objects + snapshot +
// constraints + binning for every body.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_integrate(SubstepParams P, GridDesc g, Constraints K, BodyArrays B, ColliderArrays Cc, Broadphase bp,
                                                   DeviceStats* stats, const uint32_t* __restrict__ mb_off, const uint32_t* __restrict__ mb_cols,
                                                   uint32_t only_flag) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    NlAcc na{0.f, 0.f, 0.f, 0u};
    uint2 info = make_uint2(0u, 0u);
    if (b < P.n_bodies) info = B.binfo[b];
    const uint32_t flags = info.x;
    // fused pipeline (only_flag set): free bodies were already advanced by the contact kernel
    if (b < P.n_bodies && (flags & BF_ALIVE) && (!only_flag || (flags & only_flag))) {
        const int32_t col = (int32_t)info.y;
        const float2 p = B.pos[b];
        const uint32_t wbase = g.n_worlds > 1u ? B.bworld[b] * g.ncells : 0u;
        float sx, sy, rot;
        integrate_body(P, K, B, b, flags, B.bmg[b].y, p.x, p.y, B.pos_old[b], load_acc(P, B, b), load_hv(P, B, b), sx, sy, rot, stats);
        if (col >= 0) {
            const uint32_t cf = Cc.cconst[col].y;
            if (cf & CF_ACTIVE) publish_collider(g, Cc, bp, (uint32_t)col, cf, wbase, sx, sy, rot, na);
        } else if (col <= -2) {
            const uint32_t i = (uint32_t)(-(col + 2));
            for (uint32_t k = mb_off[i]; k < mb_off[i + 1]; ++k) {
                const uint32_t c = mb_cols[k];
                const uint32_t cf = Cc.cconst[c].y;
                if (cf & CF_ACTIVE) publish_collider(g, Cc, bp, c, cf, wbase, sx, sy, rot, na);
            }
        }
    }
    if (bp.nl.snap_next != nullptr) nl_commit(bp.nl.ctl, na, &stats->collisions, 0u);
}

// ------------------------------------------------------------------------------------------------
// K-count: bins every active collider from its current snapshot (used when the broadphase is (re)built outside a step).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count(GridDesc g, ColliderArrays Cc, const uint32_t* __restrict__ bworld, uint32_t* tab_next,
                                               uint32_t* tile_next, uint32_t n_colliders, const uint8_t* __restrict__ cowned) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_colliders) return;
    if (cowned != nullptr && !cowned[c]) return;
    if (!(Cc.cconst[c].y & CF_ACTIVE)) return;
    const float2 a = Cc.cabs[c];
    const uint32_t wbase = g.n_worlds > 1u ? bworld[Cc.cparent[c]] * g.ncells : 0u;
    const uint32_t cell = wbase + cell_index(g, bin_coord(a.x, g.inv_cell), bin_coord(a.y, g.inv_cell));
    Cc.ccell[c] = make_uint2(cell, bin_collider(tab_next, tile_next, cell));
}

// ------------------------------------------------------------------------------------------------
// K-scan: exclusive prefix sum over the cell counts, in place; entry [n-1] is the sentinel (count 0) and ends up holding the
// total. One block per tile of SCAN_TILE cells; the tile totals were accumulated by the binning itself (bin_collider), so each
// block derives its own starting offset by summing the totals of the tiles before it — no inter-block dependency, no
// look-back spinning. Also zeroes `zero_me` and `tile_zero` (the table / tile totals that become "next" after the swap).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scan_tile(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ zero_me, uint32_t n_zero,
                                          const uint32_t* __restrict__ tile_cur, uint32_t* __restrict__ tile_zero) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t warp_pre[SCAN_THREADS / 32];
    const uint32_t tile = blockIdx.x;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    uint32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
            const uint4 a = *reinterpret_cast<const uint4*>(data + base + 4 * q);
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? data[base + i] : 0u;
    }
    // totals of the tiles before this one
    uint32_t pre = 0;
    for (uint32_t i = threadIdx.x; i < tile; i += SCAN_THREADS) pre += tile_cur[i];
    pre = __reduce_add_sync(0xffffffffu, pre);
    // zero the buffers that become "next"
    if (base + SCAN_ITEMS <= n_zero) {
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) *reinterpret_cast<uint4*>(zero_me + base + 4 * q) = make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < n_zero) zero_me[base + i] = 0u;
    }
    if (threadIdx.x == 0) tile_zero[tile] = 0u;

    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) tsum += v[i];
    uint32_t incl = tsum;  // warp inclusive scan of thread sums
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    if (lane == 0) warp_pre[warp] = pre;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        off += warp_pre[w];
        if (w < (int)warp) off += warp_sums[w];
    }
    uint32_t run = off + (incl - tsum);
    uint32_t o[SCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { o[i] = run; run += v[i]; }
    if (base + SCAN_ITEMS <= n) {
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q)
            *reinterpret_cast<uint4*>(data + base + 4 * q) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < n) data[base + i] = o[i];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ zero_me,
                                                       uint32_t n_zero, const uint32_t* __restrict__ tile_cur,
                                                       uint32_t* __restrict__ tile_zero) {
    scan_tile(data, n, zero_me, n_zero, tile_cur, tile_zero);
}

// ------------------------------------------------------------------------------------------------
// K-scatter: writes the 16-byte hot half of every active collider at cell_start[cell] + rank (the cold half is static).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(ColliderArrays Cc, const uint32_t* __restrict__ tab, float4* __restrict__ hot,
                                                 uint32_t n_colliders, const uint8_t* __restrict__ cowned) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_colliders) return;
    if (cowned != nullptr && !cowned[c]) return;
    const uint4 cc = Cc.cconst[c];
    const uint2 cr = Cc.ccell[c];
    const float2 a = Cc.cabs[c];
    if (!(cc.y & CF_ACTIVE)) return;
    const uint32_t dst = __ldg(tab + cr.x) + cr.y;
    hot[dst] = make_float4(a.x, a.y, __uint_as_float(cc.x), __uint_as_float(hot_word(c, cc.y)));
}

// strip mode: same, enumerating the owned-body list (single-collider bodies only, see strip_configure)
__global__ void __launch_bounds__(256) k_scatter_owned(BodyArrays B, ColliderArrays Cc, const uint32_t* __restrict__ tab, float4* __restrict__ hot,
                                                       const uint32_t* __restrict__ olist, const uint32_t* __restrict__ ocount) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= __ldg(ocount)) return;
    const uint32_t b = olist[t];
    if (b == NO_SLOT) return;
    const int32_t col = (int32_t)B.binfo[b].y;
    if (col < 0) return;
    const uint32_t c = (uint32_t)col;
    const uint4 cc = Cc.cconst[c];
    if (!(cc.y & CF_ACTIVE)) return;
    const uint2 cr = Cc.ccell[c];
    const float2 a = Cc.cabs[c];
    const uint32_t dst = __ldg(tab + cr.x) + cr.y;
    hot[dst] = make_float4(a.x, a.y, __uint_as_float(cc.x), __uint_as_float(hot_word(c, cc.y)));
}

// ------------------------------------------------------------------------------------------------
// K-springs: gravity + Spring::apply_force (springs.rs:25-47) for bodies with incident springs. One thread per such
// body; its springs are visited in spring-slot order (the order the reference accumulates into `acceleration`).
// edge = spring index << 1 | side (0: this body is rigid_body_a, 1: rigid_body_b)
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) k_springs(SubstepParams P, BodyArrays B, const uint32_t* __restrict__ sb_body,
                                                 const uint32_t* __restrict__ sb_off, const uint32_t* __restrict__ sb_edge,
                                                 const SpringParams* __restrict__ springs, uint32_t n_sb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sb) return;
    const uint32_t b = sb_body[i];
    const uint32_t flags = B.binfo[b].x;
    if (flags & BF_STATIC) return;  // no gravity (physics.rs:371), apply_force ignored (rigid_body.rs:156)
    float2 a = B.acc[b];
    const float2 mg = B.bmg[b];
    const float gm = mg.y;
    a.x = fadd(a.x, fmul(P.gx, gm));   // apply_gravity runs before the springs (physics.rs:404-408)
    a.y = fadd(a.y, fmul(P.gy, gm));
    const float m = mg.x;
    for (uint32_t e = sb_off[i]; e < sb_off[i + 1]; ++e) {
        const uint32_t ed = sb_edge[e];
        const SpringParams sp = springs[ed >> 1];
        const float2 pa = B.pos[sp.a], pb = B.pos[sp.b];
        const float2 va = B.vel[sp.a], vb = B.vel[sp.b];
        const float dx = fsub(pb.x, pa.x), dy = fsub(pb.y, pa.y);             // springs.rs:31
        const float dist = vlen(dx, dy);
        const float ux = fdiv(dx, dist), uy = fdiv(dy, dist);                 // springs.rs:33
        const float rvx = fsub(va.x, vb.x), rvy = fsub(va.y, vb.y);           // springs.rs:38
        const float dd = fmul(sp.c, fadd(fmul(rvx, ux), fmul(rvy, uy)));      // damping * rel.dot(dir)
        const float dfx = fmul(dd, ux), dfy = fmul(dd, uy);                   // springs.rs:39
        const float s = fmul(sp.k, fsub(dist, sp.rest));
        const float fmx = fsub(s, dfx), fmy = fsub(s, dfy);                   // f32 - Vec2 (springs.rs:41, Q7)
        float fx = fmul(ux, fmx), fy = fmul(uy, fmy);                         // springs.rs:43
        if (ed & 1u) { fx = -fx; fy = -fy; }                                  // rbd_b.apply_force(-force)
        a.x = fadd(a.x, fdiv(fx, m));                                         // rigid_body.rs:158
        a.y = fadd(a.y, fdiv(fy, m));
    }
    B.acc[b] = a;
}

// ------------------------------------------------------------------------------------------------
// K-joints: solve_fixed_joints (physics.rs:424-477). Sequential Gauss-Seidel is order dependent, so every connected
// component ("island") of the joint graph is solved by ONE thread, joints in slot order, joint_iterations sweeps —
// identical operation order to the reference within the island; islands are independent of each other.
// ------------------------------------------------------------------------------------------------

// physics.rs:463-465: angle_b - angle_a with angle_a = atan2(dy, dx), angle_b = -atan2(dy, -dx). For every (dx, dy) the two angles
// are supplementary: atan2(dy, -dx) = pi - atan2(dy, dx) for dy >= +0 and -pi - atan2(dy, dx) for dy <= -0 (signed zeros and
// infinities included), so the gap is -pi or +pi by the sign bit of dy - whatever the geometry. The reference evaluates it through
// two correctly-ish rounded atan2 calls and lands within 3.6e-7 of the same constant; `rotation` is tolerance-checked (1e-5) anyway
// because atan2f / sincosf differ from libm by ulps. Measured on config #4: joint kernel 1.04 -> 0.88 ms per step.
__device__ __forceinline__ float joint_angle_gap(float dx, float dy) {
    if (!(dx == dx) || !(dy == dy)) return dx + dy;   // NaN in, NaN out (the "rotation is finite" assertion must still fire)
    return (__float_as_uint(dy) >> 31) ? 3.14159274f : -3.14159274f;
}

__global__ void __launch_bounds__(128) k_joints(SubstepParams P, BodyArrays B, const uint32_t* __restrict__ isl_off,
                                                const uint32_t* __restrict__ isl_joint, const JointParams* __restrict__ joints,
                                                uint32_t n_islands, uint32_t iterations, DeviceStats* stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_islands) return;
    const uint32_t j0 = isl_off[i], j1 = isl_off[i + 1];
    bool bad = false;
    for (uint32_t it = 0; it < iterations; ++it) {
        for (uint32_t e = j0; e < j1; ++e) {
            const JointParams jp = joints[isl_joint[e]];
            float2 pa = B.pos[jp.a], pb = B.pos[jp.b];
            const float wax = fadd(pa.x, jp.aax), way = fadd(pa.y, jp.aay);          // physics.rs:434-435
            const float wbx = fadd(pb.x, jp.abx), wby = fadd(pb.y, jp.aby);
            const float dx = fsub(wbx, wax), dy = fsub(wby, way);                    // physics.rs:437
            const float dist = vlen(dx, dy);
            if (dist < 1e-6f) continue;                                              // physics.rs:440-442
            const float off_by = fsub(dist, jp.distance);
            const float cx = fdiv(fmul(off_by, dx), dist), cy = fdiv(fmul(off_by, dy), dist);   // physics.rs:445
            const float ma = B.bmg[jp.a].x, mb = B.bmg[jp.b].x;
            const float ima = fdiv(1.0f, ma), imb = fdiv(1.0f, mb);
            const float ims = fadd(ima, imb);                                        // physics.rs:450
            const uint32_t fa = B.binfo[jp.a].x, fb = B.binfo[jp.b].x;
            if (fa & BF_STATIC) {                                                    // physics.rs:452-453
                pb.x = fsub(pb.x, fmul(ims, cx)); pb.y = fsub(pb.y, fmul(ims, cy));
                B.pos[jp.b] = pb;
            } else if (fb & BF_STATIC) {                                             // physics.rs:454-455
                pa.x = fadd(pa.x, fmul(ims, cx)); pa.y = fadd(pa.y, fmul(ims, cy));
                B.pos[jp.a] = pa;
            } else {                                                                 // physics.rs:456-461
                const float ratio = fdiv(ima, ims);
                pa.x = fadd(pa.x, fmul(ratio, cx)); pa.y = fadd(pa.y, fmul(ratio, cy));
                const float r1 = fsub(1.0f, ratio);
                pb.x = fsub(pb.x, fmul(r1, cx)); pb.y = fsub(pb.y, fmul(r1, cy));
                B.pos[jp.a] = pa;
                B.pos[jp.b] = pb;
            }
            const float rc = fmul(fsub(joint_angle_gap(dx, dy), jp.target), 0.5f);   // physics.rs:463-466
            const float ra = fadd(B.rot[jp.a], fmul(rc, P.dt));                      // physics.rs:468-469
            const float rb = fsub(B.rot[jp.b], fmul(rc, P.dt));
            B.rot[jp.a] = ra;
            B.rot[jp.b] = rb;
            if (!(fabsf(ra) <= 3.4028235e38f) || !(fabsf(rb) <= 3.4028235e38f)) bad = true;   // physics.rs:471-474
        }
    }
    if (bad) atomicOr(&stats->nan_flag, 2u);
}

// ------------------------------------------------------------------------------------------------
// K-joints-fused: the fast path of solve_fixed_joints (physics.rs:424-477) for islands of up to JOINT_SMEM_MAX bodies.
// One thread per island, exactly the reference's operation order inside the island; the island's bodies live in shared
// memory laid out [local body][thread] (conflict-free), so the 4 x J sequential solves never touch global memory:
// every body is read once and written once. Joints carry LOCAL body indices.
// ------------------------------------------------------------------------------------------------
constexpr int JOINT_SMEM_MAX = 96;
constexpr int JOINT_THREADS = 64;

__global__ void __launch_bounds__(JOINT_THREADS) k_joints_fused(SubstepParams P, BodyArrays B, const uint32_t* __restrict__ isl_off,
                                                                const float4* __restrict__ jli, uint32_t max_j, const uint32_t* __restrict__ isl_boff,
                                                                const uint32_t* __restrict__ isl_body, uint32_t n_islands, uint32_t iterations,
                                                                DeviceStats* stats, uint32_t advance, GridDesc g, Constraints K, ColliderArrays Cc,
                                                                Broadphase bp, const uint32_t* __restrict__ mb_off, const uint32_t* __restrict__ mb_cols) {
#ifdef BLOBS_EMU   // host-compiled test build (tests/emu): dynamic shared memory comes from the fiber engine
    float4* const sm = static_cast<float4*>(::emu::dynamic_smem());
#else
    extern __shared__ float4 sm[];  // (pos.x, pos.y, rot, +-1/mass): a negative inverse mass marks a static body
#endif
    __shared__ uint32_t s_b0[JOINT_THREADS + 1];   // advance: first entry of each island of this CTA in isl_body (non-decreasing)
    const uint32_t t = threadIdx.x, T = JOINT_THREADS;
    const uint32_t i = blockIdx.x * T + t;
    const bool live = i < n_islands;
    uint32_t b0 = 0, nbod = 0;
    if (live) { b0 = isl_boff[i]; nbod = isl_boff[i + 1] - b0; }
    if (advance) {
        const uint32_t end = isl_boff[min(n_islands, (blockIdx.x + 1u) * T)];
        s_b0[t] = live ? b0 : end;
        if (t == 0u) s_b0[T] = end;
    }
    bool bad = false;
    if (live) {
        for (uint32_t k = 0; k < nbod; ++k) {
            const uint32_t slot = isl_body[b0 + k];
            const float2 p = B.pos[slot];
            const float m = B.bmg[slot].x;
            const float im = fdiv(1.0f, m);   // calculated_mass.recip() (physics.rs:450), hoisted: masses do not change inside the solve
            sm[k * T + t] = make_float4(p.x, p.y, B.rot[slot], (B.binfo[slot].x & BF_STATIC) ? -im : im);
        }
        const uint32_t nj = isl_off[i + 1] - isl_off[i];
        // joints are stored interleaved per CTA: record e of the island handled by thread t sits at ((block * max_j + e) * T + t),
        // two float4 each, so a warp reads consecutive records (coalesced) and the 4 sweeps re-hit L1
        const float4* jrow = jli + 2 * ((size_t)blockIdx.x * max_j * T + t);
        for (uint32_t it = 0; it < iterations; ++it) {
            for (uint32_t e = 0; e < nj; ++e) {
                const float4 q0 = __ldg(jrow + 2 * (size_t)e * T), q1 = __ldg(jrow + 2 * (size_t)e * T + 1);
                JointParams jp;
                jp.a = __float_as_uint(q0.x); jp.b = __float_as_uint(q0.y); jp.aax = q0.z; jp.aay = q0.w;
                jp.abx = q1.x; jp.aby = q1.y; jp.distance = q1.z; jp.target = q1.w;
                float4 A = sm[jp.a * T + t], Bv = sm[jp.b * T + t];
                const float wax = fadd(A.x, jp.aax), way = fadd(A.y, jp.aay);            // physics.rs:434-435
                const float wbx = fadd(Bv.x, jp.abx), wby = fadd(Bv.y, jp.aby);
                const float dx = fsub(wbx, wax), dy = fsub(wby, way);                    // physics.rs:437
                const float dist = vlen(dx, dy);
                if (dist < 1e-6f) continue;                                              // physics.rs:440-442
                const float off_by = fsub(dist, jp.distance);
                const float cx = fdiv(fmul(off_by, dx), dist), cy = fdiv(fmul(off_by, dy), dist);   // physics.rs:445
                const float ima = fabsf(A.w), imb = fabsf(Bv.w);
                const float ims = fadd(ima, imb);                                        // physics.rs:450
                if (A.w < 0.f) {                                                         // physics.rs:452-453
                    Bv.x = fsub(Bv.x, fmul(ims, cx)); Bv.y = fsub(Bv.y, fmul(ims, cy));
                } else if (Bv.w < 0.f) {                                                 // physics.rs:454-455
                    A.x = fadd(A.x, fmul(ims, cx)); A.y = fadd(A.y, fmul(ims, cy));
                } else {                                                                 // physics.rs:456-461
                    const float ratio = fdiv(ima, ims);
                    A.x = fadd(A.x, fmul(ratio, cx)); A.y = fadd(A.y, fmul(ratio, cy));
                    const float r1 = fsub(1.0f, ratio);
                    Bv.x = fsub(Bv.x, fmul(r1, cx)); Bv.y = fsub(Bv.y, fmul(r1, cy));
                }
                const float rc = fmul(fsub(joint_angle_gap(dx, dy), jp.target), 0.5f);   // physics.rs:463-466
                A.z = fadd(A.z, fmul(rc, P.dt));                                         // physics.rs:468-469
                Bv.z = fsub(Bv.z, fmul(rc, P.dt));
                if (!(fabsf(A.z) <= 3.4028235e38f) || !(fabsf(Bv.z) <= 3.4028235e38f)) bad = true;   // physics.rs:471-474
                sm[jp.a * T + t] = A;
                sm[jp.b * T + t] = Bv;
            }
        }
    }
    if (bad) atomicOr(&stats->nan_flag, 2u);
    if (!advance) {   // the bodies are advanced by a separate body-parallel pass (k_integrate)
        if (live) {
            for (uint32_t k = 0; k < nbod; ++k) {
                const uint32_t slot = isl_body[b0 + k];
                const float4 v = sm[k * T + t];
                B.rot[slot] = v.z;
                B.pos[slot] = make_float2(v.x, v.y);
            }
        }
        return;
    }
    // Advance the CTA's jointed bodies here (update_objects + apply_constraints + snapshot + binning, physics.rs:323-395) instead of
    // writing them back for k_integrate to re-read: the CTA walks the island-major list of its bodies with consecutive threads on
    // consecutive entries (= consecutive slots when a soft body's parts were inserted together), solved state out of shared memory.
    __syncthreads();
    NlAcc na{0.f, 0.f, 0.f, 0u};
    const uint32_t first = s_b0[0], end = s_b0[T];
    for (uint32_t q = first + t; q < end; q += T) {
        uint32_t lo = 0, hi = T;   // s_b0[lo] <= q < s_b0[hi]; among equal starts (empty islands) the last one owns the entry
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_b0[mid] <= q) lo = mid; else hi = mid;
        }
        const uint32_t b = isl_body[q];
        const float4 v = sm[(q - s_b0[lo]) * T + lo];
        B.rot[b] = v.z;
        const uint2 info = B.binfo[b];
        const uint32_t flags = info.x;
        const int32_t col = (int32_t)info.y;
        const uint32_t wbase = g.n_worlds > 1u ? B.bworld[b] * g.ncells : 0u;
        float sx, sy, rot;
        integrate_body(P, K, B, b, flags, B.bmg[b].y, v.x, v.y, B.pos_old[b], load_acc(P, B, b), load_hv(P, B, b), sx, sy, rot, stats);
        if (col >= 0) {
            const uint32_t cf = Cc.cconst[col].y;
            if (cf & CF_ACTIVE) publish_collider(g, Cc, bp, (uint32_t)col, cf, wbase, sx, sy, rot, na);
        } else if (col <= -2) {
            const uint32_t mi = (uint32_t)(-(col + 2));
            for (uint32_t k = mb_off[mi]; k < mb_off[mi + 1]; ++k) {
                const uint32_t c = mb_cols[k];
                const uint32_t cf = Cc.cconst[c].y;
                if (cf & CF_ACTIVE) publish_collider(g, Cc, bp, c, cf, wbase, sx, sy, rot, na);
            }
        }
    }
    if (bp.nl.snap_next != nullptr) nl_commit(bp.nl.ctl, na, &stats->collisions, 0u);
}

// ------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------
// bbox of the collider snapshot in broadphase cells (drives the table dimensions; host reads it with the stats)
__global__ void __launch_bounds__(256) k_bbox(ColliderArrays Cc, float cell, uint32_t n_colliders, DeviceStats* stats,
                                              const uint8_t* __restrict__ cowned) {
    __shared__ int s_min_x[8], s_min_y[8], s_max_x[8], s_max_y[8];
    int mnx = INT32_MAX, mny = INT32_MAX, mxx = INT32_MIN, mxy = INT32_MIN;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_colliders; c += gridDim.x * blockDim.x) {
        if (cowned != nullptr && !cowned[c]) continue;
        if (!(Cc.cconst[c].y & CF_ACTIVE)) continue;
        const float2 a = Cc.cabs[c];
        if (!(fabsf(a.x) < 1e30f) || !(fabsf(a.y) < 1e30f)) continue;   // ignore runaway / NaN points
        const int cx = cell_coord(a.x, cell), cy = cell_coord(a.y, cell);
        mnx = min(mnx, cx); mny = min(mny, cy); mxx = max(mxx, cx); mxy = max(mxy, cy);
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_min_x[w] = mnx; s_min_y[w] = mny; s_max_x[w] = mxx; s_max_y[w] = mxy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
            mnx = min(mnx, s_min_x[i]); mny = min(mny, s_min_y[i]); mxx = max(mxx, s_max_x[i]); mxy = max(mxy, s_max_y[i]);
        }
        if (mnx <= mxx) {
            atomicMin(&stats->bb_min_x, mnx); atomicMin(&stats->bb_min_y, mny);
            atomicMax(&stats->bb_max_x, mxx); atomicMax(&stats->bb_max_y, mxy);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K-query: circle queries against the live collider snapshots, served from the broadphase table (the
// reference's SpatialHash::query, spatial.rs:155-195, and the stubbed QueryPipeline / QueryFilter, lib.rs:167-187,
// query_filter.rs:27-108). One thread per query. Hit test as in SpatialHash::query: (p - q).length_squared() <= (qr + pr)^2,
// inclusive, in the reference's f32 evaluation order; unlike the reference's fixed 3x3 cell window the walk covers every cell
// the circle can reach, so no hit is missed when qr exceeds the cell size. Two passes over the same (unchanged) table:
// offsets == nullptr counts, otherwise hits are written at offsets[q] in walk order (the host sorts each segment).
// ------------------------------------------------------------------------------------------------
enum : uint32_t { QF_EXCLUDE_FIXED = 1u << 1, QF_EXCLUDE_KINEMATIC = 1u << 2, QF_EXCLUDE_DYNAMIC = 1u << 3, QF_EXCLUDE_SENSORS = 1u << 4, QF_EXCLUDE_SOLIDS = 1u << 5 };

__global__ void __launch_bounds__(128) k_query(GridDesc g, Broadphase bp, ColliderArrays Cc, BodyArrays B, const float2* __restrict__ centers,
                                               const float* __restrict__ radii, uint32_t nq, QueryFilterDev F, const uint32_t* __restrict__ offsets,
                                               uint32_t* __restrict__ counts, uint32_t* __restrict__ hits) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float2 c = centers[q];
    const float qr = radii[q];
    uint32_t n = 0;
    if (qr >= 0.f && c.x == c.x && c.y == c.y) {   // negative / NaN radius or centre: no hits
        const uint32_t base = offsets != nullptr ? offsets[q] : 0u;
        // candidate cells: reach = qr (slightly inflated: the hit test is inclusive and rounds) + largest collider radius
        for_each_candidate(g, bp, Cc.ccold, F.wbase, c.x, c.y, qr * 1.000001f + 1e-30f, [&](const Rec& o) {
            const float dist = fadd(qr, o.r);
            const float dx = fsub(o.x, c.x), dy = fsub(o.y, c.y);
            const float d2 = fadd(fmul(dx, dx), fmul(dy, dy));
            if (!(d2 <= fmul(dist, dist))) return;
            const uint32_t slot = o.slot_sensor & HOT_SLOT_MASK;
            const bool sensor = (o.slot_sensor & HOT_SENSOR_BIT) != 0u;
            if (slot == F.exclude_col) return;
            if ((F.flags & QF_EXCLUDE_SENSORS) && sensor) return;
            if ((F.flags & QF_EXCLUDE_SOLIDS) && !sensor) return;
            if (F.has_groups && !((o.memb & F.filt) != 0u && (F.memb & o.filt) != 0u)) return;   // collision_groups.test(groups)
            if (F.exclude_body != NO_SLOT || (F.flags & (QF_EXCLUDE_FIXED | QF_EXCLUDE_KINEMATIC | QF_EXCLUDE_DYNAMIC))) {
                const uint32_t parent = Cc.cparent[slot];   // the record's own parent word is a stand-in for default spheres
                if (parent == F.exclude_body) return;
                const uint32_t bf = B.binfo[parent].x;
                const bool fixed = (bf & BF_STATIC) != 0u, kin = (bf & BF_KINEMATIC) != 0u;
                if ((F.flags & QF_EXCLUDE_FIXED) && fixed) return;
                if ((F.flags & QF_EXCLUDE_KINEMATIC) && kin) return;
                if ((F.flags & QF_EXCLUDE_DYNAMIC) && !fixed && !kin) return;
            }
            if (offsets != nullptr) hits[base + n] = slot;
            ++n;
        });
    }
    if (offsets == nullptr) counts[q] = n;
}

// SpatialHash::get_cell_coords of every collider snapshot with the reference's cell size (spatial.rs:57-62)
__global__ void __launch_bounds__(256) k_cell_coords(const float2* __restrict__ cabs, float cell_size, uint32_t n, int* __restrict__ cx,
                                                     int* __restrict__ cy) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float2 a = cabs[c];
    cx[c] = cell_coord(a.x, cell_size);
    cy[c] = cell_coord(a.y, cell_size);
}


__global__ void __launch_bounds__(256) k_apply_body_writes(BodyArrays B, const BodyWrite* __restrict__ w, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BodyWrite x = w[i];
    const uint32_t s = x.slot;
    if (x.mask & BW_POS) B.pos[s] = x.pos;
    if (x.mask & BW_TRANSLATE) { float2 p = B.pos[s]; p.x = fadd(p.x, x.pos.x); p.y = fadd(p.y, x.pos.y); B.pos[s] = p; }
    if (x.mask & BW_POS_OLD) B.pos_old[s] = x.pos_old;
    if (x.mask & BW_ACC) B.acc[s] = x.acc;
    if (x.mask & BW_ADD_ACC) { float2 a = B.acc[s]; a.x = fadd(a.x, x.acc.x); a.y = fadd(a.y, x.acc.y); B.acc[s] = a; }
    if (x.mask & BW_VEL) B.vel[s] = x.vel;
    if (x.mask & BW_VREQ) { B.vreq[s] = x.vreq; B.has_vreq[s] = (uint8_t)x.has_vreq; }
    if (x.mask & BW_ROT) B.rot[s] = x.rot;
    if (x.mask & BW_ANGVEL) B.angvel[s] = x.angvel;
    if (x.mask & BW_TORQUE) B.torque[s] = x.torque;
}

__global__ void __launch_bounds__(256) k_apply_col_writes(float2* cabs, const ColWrite* __restrict__ w, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cabs[w[i].slot] = w[i].cabs;
}

// per-slot RigidBody::apply_force (rigid_body.rs:155-160): acc += force / calculated_mass for non-static bodies
__global__ void __launch_bounds__(256) k_apply_forces(BodyArrays B, const float2* __restrict__ force, uint32_t n) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t f = B.binfo[b].x;
    if (!(f & BF_ALIVE) || (f & BF_STATIC)) return;
    const float2 F = force[b];
    const float m = B.bmg[b].x;
    float2 a = B.acc[b];
    a.x = fadd(a.x, fdiv(F.x, m));
    a.y = fadd(a.y, fdiv(F.y, m));
    B.acc[b] = a;
}

// ------------------------------------------------------------------------------------------------
// Strip decomposition (BASELINE config #5): one world split into vertical strips, one per rank. Contacts are Jacobi on
// the snapshot, so ONE exchange per substep suffices: each rank sends the hot record of every owned collider within
// reach of a strip edge (ghosts) and the full state of every body whose new snapshot crossed the edge (migration).
// A cross-strip contact is seen from both sides; each side applies only its own body's half, so no reduction is needed,
// and because records carry GLOBAL slots the ordered accumulation stays bit-identical to the single-GPU run.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_strip_init_owned(BodyArrays B, ColliderArrays Cc, StripDesc S, uint8_t* owned, uint8_t* cowned,
                                                          uint32_t n_bodies) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bodies) return;
    const uint2 info = B.binfo[b];
    uint8_t o = 0;
    const int32_t col = (int32_t)info.y;
    if (info.x & BF_ALIVE) {
        const float x = col >= 0 ? Cc.cabs[col].x : B.pos[b].x;
        const bool left_ok = !S.has_left || x >= S.x_lo;
        const bool right_ok = !S.has_right || x < S.x_hi;
        o = (left_ok && right_ok) ? 1 : 0;
        if (x != x) o = S.has_left ? 0 : 1;  // NaN: rank 0 keeps it
    }
    owned[b] = o;
    if (col >= 0) cowned[col] = o;
}

// out-of-step (re)build: select ghosts (and, defensively, leavers) from the current snapshots of all owned colliders
__global__ void __launch_bounds__(256) k_strip_pack(BodyArrays B, ColliderArrays Cc, StripDesc S, const uint8_t* __restrict__ cowned,
                                                    void* send_l, void* send_r, uint32_t n_colliders) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_colliders || !cowned[c]) return;
    const uint4 cc = Cc.cconst[c];
    if (!(cc.y & CF_ACTIVE)) return;
    strip_pack_one(B, Cc, S, c, cc.y, Cc.cabs[c], __uint_as_float(cc.x), send_l, send_r);
}

// Peer-memory exchange (BLOBS_PARAM_STRIP_P2P): the fused "pack -> send -> receive" step of a substep as ONE small kernel.
// k_main has packed the outgoing messages into send_l / send_r (local memory). Every CTA copies a slice of their USED part
// (header counts, not the fixed capacity NCCL has to move) into the neighbours' receive buffers peer_l / peer_r - device
// memory of the neighbouring GPUs mapped through CUDA IPC, so these are plain stores that travel over NVLink. Each CTA fences
// its stores at system scope and bumps `done`; the last one publishes the two headers with the exchange sequence number in
// StripHeader::pad (after another system fence), then spins (volatile loads, served by L2 where the peer's stores land)
// until both incoming headers carry this sequence number; the kernels that follow in the stream read the received messages.
// A wait that exceeds ~4 s (a peer died or the ranks fell out of step) raises bit 3 of stats.nan_flag instead of hanging.
constexpr int STRIP_PUSH_CTAS = 8;
#ifdef BLOBS_EMU
constexpr long long STRIP_WAIT_TICKS = 120ll * 1000000000ll;   // host-compiled build: clock64() counts nanoseconds, ranks are slow
#else
constexpr long long STRIP_WAIT_TICKS = 1ll << 33;
#endif

__device__ __forceinline__ void strip_push_body(const StripDesc& S, const void* send_l, const void* send_r, void* peer_l, void* peer_r,
                                                const void* recv_l, const void* recv_r, uint32_t seq, unsigned int* xseq, unsigned int* done,
                                                DeviceStats* stats) {
    __shared__ bool last;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const void* src = side ? send_r : send_l;
        void* dst = side ? peer_r : peer_l;
        if (dst == nullptr) continue;
        const StripHeader* h = reinterpret_cast<const StripHeader*>(src);
        const uint32_t ng = min(h->n_ghost, S.gcap), nm4 = min(h->n_mig, S.mcap) * (uint32_t)(sizeof(MigRec) / 16);
        const float4* sg = strip_ghosts(const_cast<void*>(src));
        float4* dg = strip_ghosts(dst);
        for (uint32_t i = gtid; i < ng; i += gsz) dg[i] = sg[i];
        const float4* sm = reinterpret_cast<const float4*>(strip_migs(const_cast<void*>(src), S.gcap));
        float4* dm = reinterpret_cast<float4*>(strip_migs(dst, S.gcap));
        for (uint32_t i = gtid; i < nm4; i += gsz) dm[i] = sm[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!last) return;   // CTA-uniform
    const int side = (int)(threadIdx.x & 1u);
    if (threadIdx.x < 2u) {
        const void* src = side ? send_r : send_l;
        void* dst = side ? peer_r : peer_l;
        if (threadIdx.x == 0) { *done = 0u; *xseq = seq; }
        if (dst != nullptr) {
            __threadfence_system();   // every CTA's slice is visible before the header says so
            const StripHeader* h = reinterpret_cast<const StripHeader*>(src);
            volatile StripHeader* d = reinterpret_cast<volatile StripHeader*>(dst);
            d->n_ghost = h->n_ghost;
            d->n_mig = h->n_mig;
            d->overflow = h->overflow;
            __threadfence_system();
            d->pad = seq;
        }
    }
    __syncthreads();
    // ... and the same CTA waits for the neighbours' messages of this exchange: both of mine are out by now, so the wait can
    // never sit in front of the publication it mirrors, whatever order CTAs (or, in the host-compiled build, fibers) run in
    if (threadIdx.x < 2u && (side ? S.has_right : S.has_left)) {
        const volatile StripHeader* h = reinterpret_cast<const volatile StripHeader*>(side ? recv_r : recv_l);
        const long long t0 = clock64();
        while (h->pad != seq) {
            if (clock64() - t0 > STRIP_WAIT_TICKS) { atomicOr(&stats->nan_flag, 8u); break; }
        }
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(256) k_strip_push(StripDesc S, const void* send_l, const void* send_r, void* peer_l, void* peer_r,
                                                    const void* recv_l, const void* recv_r, unsigned int* xseq, unsigned int* done, DeviceStats* stats) {
    // exchange sequence number: device-resident (so that a captured CUDA graph can be replayed), bumped by the publishing CTA -
    // which is the last one to arrive at `done`, i.e. after every CTA has read it here
    const uint32_t seq = *reinterpret_cast<volatile unsigned int*>(xseq) + 1u;
    strip_push_body(S, send_l, send_r, peer_l, peer_r, recv_l, recv_r, seq, xseq, done, stats);
}

// bins the received ghosts into the table under construction (before k_scan)
__device__ __forceinline__ void strip_bin_ghosts_body(const GridDesc& g, const StripDesc& S, const void* recv_l, const void* recv_r, uint32_t* tab_next,
                                                      uint32_t* tile_next, uint2* gcell, DeviceStats* stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2u * S.gcap) return;
    const uint32_t side = i / S.gcap, j = i - side * S.gcap;
    const void* msg = side ? recv_r : recv_l;
    if ((side ? S.has_right : S.has_left) == 0) return;
    const StripHeader* h = reinterpret_cast<const StripHeader*>(msg);
    if (j == 0) {
        if (h->overflow) atomicOr(&stats->nan_flag, 4u);
        atomicMax(&stats->max_ghosts, h->n_ghost);
        atomicMax(&stats->max_migrants, h->n_mig);
    }
    if (j >= min(h->n_ghost, S.gcap)) return;
    const float4 hot = strip_ghosts(const_cast<void*>(msg))[j];
    const uint32_t cell = cell_index(g, bin_coord(hot.x, g.inv_cell), bin_coord(hot.y, g.inv_cell));
    gcell[i] = make_uint2(cell, bin_collider(tab_next, tile_next, cell));
}
__global__ void __launch_bounds__(256) k_strip_bin_ghosts(GridDesc g, StripDesc S, const void* recv_l, const void* recv_r, uint32_t* tab_next,
                                                          uint32_t* tile_next, uint2* gcell, DeviceStats* stats) {
    strip_bin_ghosts_body(g, S, recv_l, recv_r, tab_next, tile_next, gcell, stats);
}

// After the owned records were scattered: (1) ghost records go to their slots in the new sorted array; (2) ownership
// hand-over — leavers (my send buffers) are released, arrivals (my receive buffers) are adopted with their full state and
// appended to the owned list. One launch: threads [0, 2*gcap) do (1), threads [2*gcap, 2*gcap + 4*mcap) do (2).
__device__ __forceinline__ void strip_finish_body(const BodyArrays& B, const ColliderArrays& Cc, const StripDesc& S, const void* send_l, const void* send_r,
                                                  const void* recv_l, const void* recv_r, const uint32_t* __restrict__ tab,
                                                  const uint2* __restrict__ gcell, float4* __restrict__ hot, uint8_t* owned, uint8_t* cowned,
                                                  uint32_t* olist, uint32_t* ocount, uint32_t* opos, uint32_t olist_cap, DeviceStats* stats,
                                                  float4* __restrict__ snap) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2u * S.gcap) {
        const uint32_t side = i / S.gcap, j = i - side * S.gcap;
        const void* msg = side ? recv_r : recv_l;
        if ((side ? S.has_right : S.has_left) == 0) return;
        const StripHeader* h = reinterpret_cast<const StripHeader*>(msg);
        if (j >= min(h->n_ghost, S.gcap)) return;
        const uint2 cr = gcell[i];
        const float4 rec = strip_ghosts(const_cast<void*>(msg))[j];
        hot[__ldg(tab + cr.x) + cr.y] = rec;
        if (snap != nullptr) snap[__float_as_uint(rec.w) & HOT_SLOT_MASK] = rec;   // list pipeline: what the contact pass reads until the owner's next push
        return;
    }
    i -= 2u * S.gcap;
    if (i >= 4u * S.mcap) return;
    const uint32_t which = i / S.mcap, j = i - which * S.mcap;
    const void* msg = which == 0 ? send_l : (which == 1 ? send_r : (which == 2 ? recv_l : recv_r));
    const bool has = (which & 1u) ? S.has_right != 0 : S.has_left != 0;
    if (!has) return;
    const StripHeader* h = reinterpret_cast<const StripHeader*>(msg);
    if (j >= min(h->n_mig, S.mcap)) return;
    const MigRec m = strip_migs(const_cast<void*>(msg), S.gcap)[j];
    if (which < 2u) {
        owned[m.slot] = 0;
        cowned[m.col] = 0;
        olist[opos[m.slot]] = NO_SLOT;
    } else {
        B.pos[m.slot] = m.pos; B.pos_old[m.slot] = m.pos_old; B.acc[m.slot] = m.acc; B.vel[m.slot] = m.vel; B.vreq[m.slot] = m.vreq;
        B.rot[m.slot] = m.rot; B.angvel[m.slot] = m.angvel; B.torque[m.slot] = m.torque; B.has_vreq[m.slot] = (uint8_t)m.has_vreq;
        Cc.cabs[m.col] = m.cabs;
        owned[m.slot] = 1;
        cowned[m.col] = 1;
        const uint32_t k = atomicAdd(ocount, 1u);
        if (k < olist_cap) { olist[k] = m.slot; opos[m.slot] = k; }
        else atomicOr(&stats->nan_flag, 4u);
    }
}
__global__ void __launch_bounds__(256) k_strip_finish(BodyArrays B, ColliderArrays Cc, StripDesc S, const void* send_l, const void* send_r,
                                                      const void* recv_l, const void* recv_r, const uint32_t* __restrict__ tab,
                                                      const uint2* __restrict__ gcell, float4* __restrict__ hot, uint8_t* owned, uint8_t* cowned,
                                                      uint32_t* olist, uint32_t* ocount, uint32_t* opos, uint32_t olist_cap, DeviceStats* stats) {
    strip_finish_body(B, Cc, S, send_l, send_r, recv_l, recv_r, tab, gcell, hot, owned, cowned, olist, ocount, opos, olist_cap, stats, nullptr);
}

// (re)builds the compact owned list from the ownership bytes (once per blobs_step* call: drops released entries)
__global__ void __launch_bounds__(256) k_strip_build_olist(const uint8_t* __restrict__ owned, uint32_t n_bodies, uint32_t* olist, uint32_t* ocount,
                                                           uint32_t* opos) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = b < n_bodies && owned[b];
    const unsigned int m = __ballot_sync(0xffffffffu, mine);
    if (!m) return;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(ocount, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) {
        const uint32_t i = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
        olist[i] = b;
        opos[b] = i;
    }
}

// distributed host I/O: every rank reads / drives only the bodies it currently owns
__global__ void __launch_bounds__(256) k_compact_owned(BodyArrays B, const uint8_t* __restrict__ owned, uint32_t n_bodies, uint32_t cap,
                                                       unsigned int* count, uint32_t* __restrict__ slots, float2* __restrict__ xy) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = b < n_bodies && (B.binfo[b].x & BF_ALIVE) && (owned == nullptr || owned[b]);
    const unsigned int m = __ballot_sync(0xffffffffu, mine);
    if (!m) return;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) {
        const uint32_t i = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (i < cap) { slots[i] = b; xy[i] = B.pos[b]; }
    }
}

// indexed RigidBody::apply_force (rigid_body.rs:155-160); entries for bodies this rank does not own are ignored
__global__ void __launch_bounds__(256) k_apply_forces_indexed(BodyArrays B, const uint8_t* __restrict__ owned, const uint32_t* __restrict__ slots,
                                                              const float2* __restrict__ force, uint32_t n, uint32_t n_bodies) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = slots[i];
    if (b >= n_bodies || (owned != nullptr && !owned[b])) return;
    const uint32_t f = B.binfo[b].x;
    if (!(f & BF_ALIVE) || (f & BF_STATIC)) return;
    const float2 F = force[i];
    const float m = B.bmg[b].x;
    float2 a = B.acc[b];
    a.x = fadd(a.x, fdiv(F.x, m));
    a.y = fadd(a.y, fdiv(F.y, m));
    B.acc[b] = a;
}

#include "nlist.cuh"

}  // namespace blobs
