// Probe (round 2): can a kernel inside a captured CUDA graph tail-launch the rebuild kernels itself (CUDA dynamic parallelism,
// cudaStreamTailLaunch), is the stream successor held back until the tail launches have run, and what does a substep save against
// four self-gating launches?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -o cdp_probe cdp_probe.cu -lcudadevrt
#include <cuda_runtime.h>
#include <cstdio>
__global__ void child(int* p, int v) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, v); }
__global__ void gated(const int* need, int* p) { if (*need == 0) return; if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1); }
__global__ void work(float* x, int n, const int* need, int* p, int tail) {   // stands in for k_step: ~1 M threads, a little memory traffic
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] * 1.0001f + 1.f;
    if (tail && threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && *need) {
        child<<<1184, 256, 0, cudaStreamTailLaunch>>>(p, 1);
        child<<<1184, 256, 0, cudaStreamTailLaunch>>>(p, 10);
        child<<<1184, 256, 0, cudaStreamTailLaunch>>>(p, 100);
        child<<<1184, 256, 0, cudaStreamTailLaunch>>>(p, 1000);
    }
}
__global__ void after(int* p, int* out) { *out = *p; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    int *p, *o, *need; float* x; const int n = 1 << 20;
    CK(cudaMalloc(&p, 4)); CK(cudaMalloc(&o, 4)); CK(cudaMalloc(&need, 4)); CK(cudaMalloc(&x, n * 4));
    CK(cudaMemset(p, 0, 4)); CK(cudaMemset(x, 0, n * 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int variant = 0; variant < 2; ++variant) {        // 0: four self-gating launches per substep, 1: tail launch from the work kernel
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int sub = 0; sub < 8; ++sub) {
            if (variant == 0) for (int k = 0; k < 4; ++k) gated<<<1184, 256, 0, s>>>(need, p);
            work<<<n / 256, 256, 0, s>>>(x, n, need, p, variant);
        }
        after<<<1, 1, 0, s>>>(p, o);
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int nd = 0; nd < 2; ++nd) {
            CK(cudaMemset(need, 0, 4)); CK(cudaMemset(p, 0, 4));
            if (nd) { int one = 1; CK(cudaMemcpy(need, &one, 4, cudaMemcpyHostToDevice)); }
            for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, s));
            CK(cudaStreamSynchronize(s));
            CK(cudaMemset(p, 0, 4));
            cudaEventRecord(a, s);
            for (int i = 0; i < 200; ++i) CK(cudaGraphLaunch(ge, s));
            cudaEventRecord(b, s);
            CK(cudaStreamSynchronize(s));
            float ms = 0; cudaEventElapsedTime(&ms, a, b);
            int h = -1; CK(cudaMemcpy(&h, o, 4, cudaMemcpyDeviceToHost));
            printf("variant %d (%s) need=%d: %.2f us per substep, counter seen by the successor = %d (want %d)\n", variant,
                   variant ? "tail launch" : "4 gated launches", nd, ms * 1e3 / (200 * 8), h, nd ? (variant ? 200 * 8 * 1111 : 200 * 8 * 4) : 0);
        }
    }
    return 0;
}
