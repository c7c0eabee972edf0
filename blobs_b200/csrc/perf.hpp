// Host-side instrumentation that the reference's callers see each frame:
//  * the process-global perf-counter registry (reference blobs/src/perf_counters.rs:3-87): name -> (count, decayed average);
//    Physics::step feeds "collisions" (physics.rs:316), the demo calls perf_counters_new_frame(delta) once per frame
//    (demo/src/main.rs:223) and lists every counter in its perf panel (main.rs:291-300);
//  * profiler ranges named like the reference's tracy spans (physics.rs:79,92,242,324,398,402) - emitted as NVTX ranges so an
//    Nsight Systems timeline of a game using this library reads like a Tracy capture of the reference.
// No device code here; the registry is a plain std::map behind a mutex (the reference uses a global AtomicRefCell).
#pragma once
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>

#if !defined(BLOBS_EMU) && !defined(BLOBS_NO_NVTX)
#include <nvtx3/nvToolsExt.h>   // header-only: a no-op function-pointer check unless a profiler injected itself
#define BLOBS_HAVE_NVTX 1
#endif

namespace blobs {

struct PerfCounter {       // perf_counters.rs:11-15
    uint64_t count = 0;
    double decayed_average = 0.0;
};

class PerfCounters {       // perf_counters.rs:6-9,17-50
   public:
    static PerfCounters& global() {
        static PerfCounters g;
        return g;
    }
    void update(const char* name, uint64_t count) {             // update_counter, perf_counters.rs:22-25
        std::lock_guard<std::mutex> l(mu);
        counters[name].count = count;
    }
    void inc(const char* name, uint64_t by) {                   // perf_counter_inc, perf_counters.rs:71-76 (creates the counter at 0 + by)
        std::lock_guard<std::mutex> l(mu);
        counters[name].count += by;
    }
    void new_frame(double delta) {                              // new_frame, perf_counters.rs:27-33
        std::lock_guard<std::mutex> l(mu);
        for (auto& kv : counters) {
            PerfCounter& c = kv.second;
            c.decayed_average = c.decayed_average * (1.0 - delta) + (double)c.count * delta;
            c.count = 0;
        }
    }
    PerfCounter get(const char* name) {                         // get_counter, perf_counters.rs:35-41: (0, 0.0) when absent
        std::lock_guard<std::mutex> l(mu);
        auto it = counters.find(name);
        return it == counters.end() ? PerfCounter{} : it->second;
    }
    void reset() {                                              // reset_counters, perf_counters.rs:43-45
        std::lock_guard<std::mutex> l(mu);
        counters.clear();
    }
    size_t size() {
        std::lock_guard<std::mutex> l(mu);
        return counters.size();
    }
    // i-th counter in name order (the reference iterates a HashMap, i.e. in no particular order)
    bool at(size_t i, std::string* name, PerfCounter* out) {
        std::lock_guard<std::mutex> l(mu);
        if (i >= counters.size()) return false;
        auto it = counters.begin();
        std::advance(it, (long)i);
        *name = it->first;
        *out = it->second;
        return true;
    }

   private:
    std::mutex mu;
    std::map<std::string, PerfCounter> counters;
};

// events.rs:22-40,62-64: process-global history of soft errors, capped at 1000 entries (oldest dropped). The reference only
// ever pushes two messages: "removing a non-existent rigid body" (rigid_body.rs:266-275, Severity::Error) and
// "rbd removed because colliders.len() == 0" (collider.rs:143-158, Severity::Info).
struct PhysicsEventRec {
    double real_time = 0.0, unpaused_time = 0.0;   // TimeData (events.rs:5-18): the reference never advances it
    float px = 0.f, py = 0.f;
    bool has_position = false;
    int severity = 0;                              // Severity (events.rs:52-60): Trace=0 .. Critical=5
    uint64_t col_handle = 0, rbd_handle = 0;       // 0 = None
    std::string message;
};

class EventHistory {
   public:
    static constexpr size_t CAP = 1000;            // events.rs:35-38
    static EventHistory& global() {
        static EventHistory g;
        return g;
    }
    void push(PhysicsEventRec e) {
        std::lock_guard<std::mutex> l(mu);
        events.push_back(std::move(e));
        while (events.size() > CAP) events.pop_front();
    }
    size_t size() {
        std::lock_guard<std::mutex> l(mu);
        return events.size();
    }
    bool at(size_t i, PhysicsEventRec* out) {
        std::lock_guard<std::mutex> l(mu);
        if (i >= events.size()) return false;
        *out = events[i];
        return true;
    }
    void clear() {
        std::lock_guard<std::mutex> l(mu);
        events.clear();
    }

   private:
    std::mutex mu;
    std::deque<PhysicsEventRec> events;
};

// RAII profiler range, the counterpart of `let _span = tracy_span!("name")` (lib.rs:195-209)
struct Span {
#ifdef BLOBS_HAVE_NVTX
    explicit Span(const char* name) { nvtxRangePushA(name); }
    ~Span() { nvtxRangePop(); }
#else
    explicit Span(const char*) {}
#endif
    Span(const Span&) = delete;
    Span& operator=(const Span&) = delete;
};

}  // namespace blobs
