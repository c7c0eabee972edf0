// TEST INFRASTRUCTURE: the eight NCCL entry points world.cu binds, implemented over Unix-domain sockets between PROCESSES
// on one machine, for the host-compiled test build (tests/emu). It lets the strip-decomposed data path (ghost / migration
// exchange with both neighbours every substep) run as two or more CPU processes and be compared bit-for-bit with the
// single-world result. Buffers are host memory there, "streams" are synchronous: a grouped send/recv set is executed at
// ncclGroupEnd (sends on a helper thread, receives on the caller) so that neighbours exchanging simultaneously cannot block
// each other.
#include <nccl.h>

#include <errno.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

struct ncclComm {
    int rank = 0, nranks = 0;
    std::string prefix;
    int listen_fd = -1;
    std::vector<int> peer;   // connected socket per peer rank (-1 = none)
};

namespace {
struct Op { bool send; void* buf; size_t n; int peer; ncclComm* comm; };
thread_local int g_depth = 0;
thread_local std::vector<Op> g_ops;

bool write_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        ssize_t k = ::write(fd, c, n);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        c += k; n -= (size_t)k;
    }
    return true;
}
bool read_all(int fd, void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n) {
        ssize_t k = ::read(fd, c, n);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        if (k == 0) return false;
        c += k; n -= (size_t)k;
    }
    return true;
}
sockaddr_un addr_of(const std::string& prefix, int rank) {
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    snprintf(a.sun_path, sizeof(a.sun_path), "%s_%d", prefix.c_str(), rank);
    return a;
}
ncclResult_t run_ops(std::vector<Op>& ops) {
    bool ok_send = true, ok_recv = true;
    std::thread sender([&] {
        for (const Op& o : ops)
            if (o.send) ok_send = ok_send && o.comm->peer[o.peer] >= 0 && write_all(o.comm->peer[o.peer], o.buf, o.n);
    });
    for (const Op& o : ops)
        if (!o.send) ok_recv = ok_recv && o.comm->peer[o.peer] >= 0 && read_all(o.comm->peer[o.peer], o.buf, o.n);
    sender.join();
    ops.clear();
    return ok_send && ok_recv ? ncclSuccess : ncclUnhandledCudaError;
}
}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    memset(id->internal, 0, sizeof(id->internal));
    std::random_device rd;
    snprintf(id->internal, sizeof(id->internal), "/tmp/blobs_fake_nccl_%08x%08x_%d", rd(), rd(), (int)getpid());
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* out, int nranks, ncclUniqueId id, int rank) {
    ncclComm* c = new ncclComm();
    c->rank = rank; c->nranks = nranks; c->prefix = id.internal;
    c->peer.assign(nranks, -1);
    c->listen_fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un me = addr_of(c->prefix, rank);
    ::unlink(me.sun_path);
    if (c->listen_fd < 0 || ::bind(c->listen_fd, (sockaddr*)&me, sizeof(me)) != 0 || ::listen(c->listen_fd, nranks) != 0) return ncclUnhandledCudaError;
    // full mesh: connect to every lower rank, accept from every higher one
    for (int p = 0; p < rank; ++p) {
        sockaddr_un a = addr_of(c->prefix, p);
        int fd = -1;
        for (int tries = 0; tries < 3000; ++tries) {   // the peer may not be listening yet
            fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
            if (::connect(fd, (sockaddr*)&a, sizeof(a)) == 0) break;
            ::close(fd); fd = -1;
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        if (fd < 0 || !write_all(fd, &rank, sizeof(rank))) return ncclUnhandledCudaError;
        c->peer[p] = fd;
    }
    for (int k = rank + 1; k < nranks; ++k) {
        int fd = ::accept(c->listen_fd, nullptr, nullptr);
        int who = -1;
        if (fd < 0 || !read_all(fd, &who, sizeof(who)) || who <= rank || who >= nranks) return ncclUnhandledCudaError;
        c->peer[who] = fd;
    }
    *out = c;
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) {
    if (!c) return ncclSuccess;
    for (int fd : c->peer) if (fd >= 0) ::close(fd);
    if (c->listen_fd >= 0) ::close(c->listen_fd);
    sockaddr_un me = addr_of(c->prefix, c->rank);
    ::unlink(me.sun_path);
    delete c;
    return ncclSuccess;
}

ncclResult_t ncclGroupStart() { g_depth++; return ncclSuccess; }
ncclResult_t ncclGroupEnd() {
    if (--g_depth > 0) return ncclSuccess;
    return run_ops(g_ops);
}
ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t, int peer, ncclComm_t comm, cudaStream_t) {
    g_ops.push_back(Op{true, const_cast<void*>(buf), count, peer, comm});
    return g_depth ? ncclSuccess : run_ops(g_ops);
}
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t, int peer, ncclComm_t comm, cudaStream_t) {
    g_ops.push_back(Op{false, buf, count, peer, comm});
    return g_depth ? ncclSuccess : run_ops(g_ops);
}
const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL: socket exchange failed"; }

}  // extern "C"
