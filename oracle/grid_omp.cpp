// TEST / BENCH INFRASTRUCTURE — never linked into or called by the product (libblobs_b200.so).
//
// "CPU cell-list restatement": Physics::step (physics.rs:397-422) for the sphere scenes of BASELINE configs #1/#2/#3/#5
// (every body dynamic, ONE zero-offset collider per body with collider slot == body slot, groups ALL, no sensors, no springs,
// no joints), restated so that it can use all host cores. NOT the reference algorithm: the reference's live collision path is
// the O(C^2) all-pairs loop (physics.rs:241-317) on one thread. What is kept exactly is the ARITHMETIC and its ORDER:
//   * contacts are detected and measured on the collider snapshots of the previous substep (physics.rs:264-267,360-366);
//   * a body's position receives its contributions in the pair loop's order, which for a single-collider body is ascending
//     partner slot, each computed as in physics.rs:291-300 / 319-321;
//   * update_objects, the old_dt quirk (physics.rs:338-339), velocity_request, apply_constraints as in blobs_oracle.cpp.
// So the results are bit-identical to oracle/blobs_oracle.cpp (brute force or its sequential grid variant) as long as no pair is
// closer than 1e-6 (physics.rs:272-286 is inherently sequential; such a pair is counted and reported, the caller must treat the
// run as unsupported). tests/test_grid_omp.py pins that equality on cfg1, a falling lattice and a dense pile.
// Two uses: (1) the labelled all-cores CPU number beside the GPU's (bench.py: cpu_baseline.grid_restatement); (2) a checker that
// finishes 1M spheres x hundreds of steps, for the full-size parity tests.
// Build: g++ -O2 -fopenmp -ffp-contract=off -fno-fast-math (see Makefile).
#include <omp.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct V2 { float x, y; };
inline float len2f(float x, float y) { return std::sqrt(x * x + y * y); }   // glam Vec2::length

struct Contrib { uint32_t j; float cx, cy; };

}  // namespace

extern "C" {

// All arrays have n entries (xy arrays 2n floats, interleaved). State arrays are updated in place.
// constraints: n_con x (x, y, radius). old_dt: in/out (Physics::old_dt). Returns 0, or 1 if n == 0.
int blobs_grid_omp_step(uint64_t n64, float* pos, float* pos_old, float* acc, float* vel, float* snap, const float* radius, const float* mass,
                        const float* gravity_mod, float* vreq, uint8_t* has_vreq, float gx, float gy, const float* constraints, int n_con,
                        float* old_dt_io, double delta, uint32_t substeps, uint32_t steps, int threads, uint64_t* collisions_out,
                        uint64_t* coincident_out, double* seconds_out) {
    const uint32_t n = (uint32_t)n64;
    if (n == 0) return 1;
    if (threads > 0) omp_set_num_threads(threads);
    V2* P = reinterpret_cast<V2*>(pos);
    V2* PO = reinterpret_cast<V2*>(pos_old);
    V2* A = reinterpret_cast<V2*>(acc);
    V2* VEL = reinterpret_cast<V2*>(vel);
    V2* S = reinterpret_cast<V2*>(snap);
    V2* VR = reinterpret_cast<V2*>(vreq);
    float old_dt = *old_dt_io;
    float rmax = 0.f;
    for (uint32_t i = 0; i < n; ++i) rmax = std::max(rmax, radius[i]);
    const float cs = std::max(2.0f * rmax * 1.001f, 1e-3f);
    uint64_t collisions = 0, coincident = 0;
    std::vector<uint32_t> cell(n), rank(n), start, sorted(n);
    const double t0 = omp_get_wtime();
    for (uint32_t st = 0; st < steps; ++st) {
        const float dt = (float)delta / (float)substeps;   // physics.rs:399 (delta is f32 there: step(delta as f32))
        for (uint32_t sub = 0; sub < substeps; ++sub) {
            // ---- cell list over the snapshots (counting sort; order inside a cell is irrelevant: contributions are sorted by slot)
            float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
#pragma omp parallel for reduction(min : mnx, mny) reduction(max : mxx, mxy) schedule(static)
            for (uint32_t i = 0; i < n; ++i) {
                const V2 s = S[i];
                if (std::isfinite(s.x) && std::isfinite(s.y)) {
                    mnx = std::min(mnx, s.x); mny = std::min(mny, s.y);
                    mxx = std::max(mxx, s.x); mxy = std::max(mxy, s.y);
                }
            }
            if (!(mnx <= mxx)) { mnx = mny = 0.f; mxx = mxy = 0.f; }
            const int64_t cx0 = (int64_t)std::floor(mnx / cs) - 1, cy0 = (int64_t)std::floor(mny / cs) - 1;
            const int64_t W = (int64_t)std::floor(mxx / cs) - cx0 + 2, H = (int64_t)std::floor(mxy / cs) - cy0 + 2;
            if (W * H > (int64_t)1 << 31) return 2;
            const uint32_t ncell = (uint32_t)(W * H);
            start.assign((size_t)ncell + 1, 0u);
            auto cell_of = [&](V2 s, int64_t& cx, int64_t& cy) {
                cx = (int64_t)std::floor(s.x / cs) - cx0;
                cy = (int64_t)std::floor(s.y / cs) - cy0;
                if (!(cx >= 0 && cx < W)) cx = 0;   // non-finite snapshot: parked in a corner cell (it contacts nothing: NaN < x is false)
                if (!(cy >= 0 && cy < H)) cy = 0;
            };
#pragma omp parallel for schedule(static)
            for (uint32_t i = 0; i < n; ++i) {
                int64_t cx, cy;
                cell_of(S[i], cx, cy);
                const uint32_t c = (uint32_t)(cy * W + cx);
                cell[i] = c;
                uint32_t r;
#pragma omp atomic capture
                r = start[c + 1]++;
                rank[i] = r;
            }
            for (uint32_t c = 0; c < ncell; ++c) start[c + 1] += start[c];   // (sequential scan: ~1 ms per million cells)
#pragma omp parallel for schedule(static)
            for (uint32_t i = 0; i < n; ++i) sorted[start[cell[i]] + rank[i]] = i;

            // ---- contacts: gather per body, contributions in ascending partner slot (physics.rs:248-300)
            uint64_t ncol = 0, ncoin = 0;
#pragma omp parallel for reduction(+ : ncol, ncoin) schedule(dynamic, 1024)
            for (uint32_t i = 0; i < n; ++i) {
                const V2 si = S[i];
                const float ri = radius[i], mi = mass[i];
                Contrib buf[64];
                std::vector<Contrib> big;
                uint32_t nc = 0;
                int64_t cx, cy;
                cell_of(si, cx, cy);
                for (int64_t yy = std::max<int64_t>(cy - 2, 0); yy <= std::min<int64_t>(cy + 2, H - 1); ++yy)      // 5x5: one spare ring, as in
                    for (int64_t xx = std::max<int64_t>(cx - 2, 0); xx <= std::min<int64_t>(cx + 2, W - 1); ++xx) {  // Physics::grid_collisions
                        const uint32_t c = (uint32_t)(yy * W + xx);
                        for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
                            const uint32_t j = sorted[k];
                            if (j == i) continue;
                            const bool i_am_a = i > j;   // col_a = later slot (physics.rs:248-249)
                            const V2 sj = S[j];
                            const float ax = i_am_a ? si.x - sj.x : sj.x - si.x, ay = i_am_a ? si.y - sj.y : sj.y - si.y;   // :264
                            const float dist = len2f(ax, ay);
                            const float min_dist = i_am_a ? ri + radius[j] : radius[j] + ri;                               // :267
                            if (!(dist < min_dist)) continue;                                                               // :269
                            if (dist < 1e-6f) { ncoin++; continue; }   // sequential branch (:272-286): unsupported here, reported
                            const float nx = ax / dist, ny = ay / dist;                                                     // :292
                            const float dl = min_dist - dist;                                                               // :294
                            const float ma = i_am_a ? mi : mass[j], mb = i_am_a ? mass[j] : mi;
                            const float ratio = 1.0f - ma / (ma + mb);                                                      // :319-321
                            Contrib c2;
                            c2.j = j;
                            if (i_am_a) { const float kk = ratio * dl; c2.cx = kk * nx; c2.cy = kk * ny; }                  // :298
                            else { const float kk = (1.0f - ratio) * dl; c2.cx = -(kk * nx); c2.cy = -(kk * ny); }          // :299
                            if (i_am_a) ncol++;
                            if (nc < 64) buf[nc] = c2; else { if (nc == 64) big.assign(buf, buf + 64); big.push_back(c2); }
                            nc++;
                        }
                    }
                Contrib* cb = nc <= 64 ? buf : big.data();
                std::sort(cb, cb + nc, [](const Contrib& a, const Contrib& b) { return a.j < b.j; });
                V2 p = P[i];
                for (uint32_t k = 0; k < nc; ++k) { p.x += cb[k].cx; p.y += cb[k].cy; }
                P[i] = p;
            }
            collisions += ncol;
            coincident += ncoin;

            // ---- gravity + update_objects (physics.rs:323-358,369-375); the first non-static body sees dt / old_dt (Q2)
            const float ratio_first = dt / old_dt, ratio_rest = dt / dt;
#pragma omp parallel for schedule(static)
            for (uint32_t i = 0; i < n; ++i) {
                V2 a = A[i];
                a.x += gx * gravity_mod[i];
                a.y += gy * gravity_mod[i];
                V2 p = P[i], po = PO[i];
                if (has_vreq[i]) {
                    po.x = p.x - VR[i].x * dt;
                    po.y = p.y - VR[i].y * dt;
                    has_vreq[i] = 0;
                }
                const float ratio = i == 0 ? ratio_first : ratio_rest;
                const float dx = (p.x - po.x) * ratio, dy = (p.y - po.y) * ratio;
                PO[i] = p;
                p.x += dx + a.x * dt * dt;
                p.y += dy + a.y * dt * dt;
                A[i] = V2{0.f, 0.f};
                VEL[i] = V2{dx / dt, dy / dt};
                S[i] = p;   // snapshot = from_angle_translation(0, position) * identity offset: taken BEFORE the clamp (physics.rs:360-366,419-420)
                for (int c = 0; c < n_con; ++c) {   // physics.rs:377-395
                    const float ox = constraints[3 * c], oy = constraints[3 * c + 1], R = constraints[3 * c + 2];
                    const float tx = p.x - ox, ty = p.y - oy;
                    const float d = len2f(tx, ty);
                    if (d > R) { p.x = ox + (tx / d) * R; p.y = oy + (ty / d) * R; }
                }
                P[i] = p;
            }
            old_dt = dt;
        }
    }
    *seconds_out = omp_get_wtime() - t0;
    *old_dt_io = old_dt;
    *collisions_out = collisions;
    *coincident_out = coincident;
    return 0;
}

int blobs_grid_omp_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
