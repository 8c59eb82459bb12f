// chain_probe.cu -- what does a copy-in -> kernel -> copy-out chain cost on this box, independent of the ISMPC kernels?
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o chain_probe chain_probe.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void spin_kernel(const char* in, char* out, long long ns)
{
    // every CTA waits `ns` nanoseconds (stands in for a latency-bound tick), thread 0 of each CTA moves one 128-byte record
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long t = t0;
    while (t - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (threadIdx.x < 8) reinterpret_cast<double2*>(out)[blockIdx.x * 8 + threadIdx.x] = reinterpret_cast<const double2*>(in)[blockIdx.x * 8 + threadIdx.x];
}

// the tick's real access pattern: per CTA 72 B of state[], 24 B of walk[], 40 B of inst[] (three arrays), 8 bytes per lane
__global__ void spin3_kernel(const double* state, const double* walk, const double* inst, char* out, long long ns)
{
    const int i = blockIdx.x, l = threadIdx.x;
    double v = 0.0;
    if (l < 9) v = state[9 * i + l];
    else if (l < 12) v = walk[3 * i + (l - 9)];
    else if (l < 17) v = inst[5 * i + (l - 12)];
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long t = t0;
    while (t - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (l < 16) reinterpret_cast<double*>(out)[blockIdx.x * 16 + l] = v;
}
// the copy engine's job done by SMs: coalesced 16-byte loads from pinned host memory, stores to device memory
__global__ void gather_kernel(const int4* src, int4* dst, int n16)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) dst[i] = src[i];
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
    const size_t IN = 139264, OUT = 131072;
    const int S = 8, K = 2000;
    std::vector<cudaStream_t> st(S);
    std::vector<char*> hin(S), hout(S), din(S), dout(S);
    for (int s = 0; s < S; ++s) {
        cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
        cudaHostAlloc(&hin[s], IN, cudaHostAllocPortable); cudaHostAlloc(&hout[s], OUT, cudaHostAllocPortable);
        cudaMalloc(&din[s], IN); cudaMalloc(&dout[s], OUT);
    }
    auto report = [&](const char* name, double t0, double t1, double t2, int k) {
        printf("%-64s %7.2f us per step (host enqueue %.2f us)\n", name, (t2 - t0) / k * 1e6, (t1 - t0) / k * 1e6);
    };
    for (int rep = 0; rep < 2; ++rep) {
        cudaDeviceSynchronize();
        double t0 = now();
        for (int k = 0; k < K; ++k) cudaMemcpyAsync(din[0], hin[0], IN, cudaMemcpyHostToDevice, st[0]);
        double t1 = now(); cudaDeviceSynchronize(); double t2 = now();
        if (rep) report("H2D 139264 B back to back, 1 stream", t0, t1, t2, K);
        t0 = now();
        for (int k = 0; k < K; ++k) cudaMemcpyAsync(hout[0], dout[0], OUT, cudaMemcpyDeviceToHost, st[0]);
        t1 = now(); cudaDeviceSynchronize(); t2 = now();
        if (rep) report("D2H 131072 B back to back, 1 stream", t0, t1, t2, K);
        t0 = now();
        for (int k = 0; k < K; ++k) { cudaMemcpyAsync(din[k % S], hin[k % S], IN, cudaMemcpyHostToDevice, st[k % S]); cudaMemcpyAsync(hout[k % S], dout[k % S], OUT, cudaMemcpyDeviceToHost, st[k % S]); }
        t1 = now(); cudaDeviceSynchronize(); t2 = now();
        if (rep) report("H2D + D2H per step, 8 streams round robin", t0, t1, t2, K);
        t0 = now();
        for (int k = 0; k < K; ++k) spin_kernel<<<1024, 64, 0, st[k % S]>>>(din[k % S], dout[k % S], 0);
        t1 = now(); cudaDeviceSynchronize(); t2 = now();
        if (rep) report("empty kernel 1024x64, 8 streams round robin", t0, t1, t2, K);
        for (long long ns : {0LL, 5000LL, 8000LL}) {
            t0 = now();
            for (int k = 0; k < K; ++k) {
                const int s = k % S;
                cudaMemcpyAsync(din[s], hin[s], IN, cudaMemcpyHostToDevice, st[s]);
                spin_kernel<<<1024, 64, 0, st[s]>>>(din[s], dout[s], ns);
                cudaMemcpyAsync(hout[s], dout[s], OUT, cudaMemcpyDeviceToHost, st[s]);
            }
            t1 = now(); cudaDeviceSynchronize(); t2 = now();
            char nm[128]; snprintf(nm, sizeof nm, "H2D -> kernel(%lld ns per CTA) -> D2H, 8 streams round robin", ns);
            if (rep) report(nm, t0, t1, t2, K);
            t0 = now();
            for (int k = 0; k < K; ++k) {
                const int s = k % S;
                cudaMemcpyAsync(din[s], hin[s], IN, cudaMemcpyHostToDevice, st[s]);
                spin_kernel<<<1024, 64, 0, st[s]>>>(din[s], hout[s], ns);
            }
            t1 = now(); cudaDeviceSynchronize(); t2 = now();
            snprintf(nm, sizeof nm, "H2D -> kernel(%lld ns) writing pinned host memory, 8 streams", ns);
            if (rep) report(nm, t0, t1, t2, K);
            t0 = now();
            for (int k = 0; k < K; ++k) {
                const int s = k % S;
                const double* b = reinterpret_cast<const double*>(hin[s]);
                spin3_kernel<<<1024, 64, 0, st[s]>>>(b, b + 9 * 1024, b + 12 * 1024, hout[s], ns);
            }
            t1 = now(); cudaDeviceSynchronize(); t2 = now();
            snprintf(nm, sizeof nm, "kernel(%lld ns) reading 3 host arrays (72/24/40 B per CTA), writing host, 8 streams", ns);
            if (rep) report(nm, t0, t1, t2, K);
            t0 = now();
            for (int k = 0; k < K; ++k) {
                const int s = k % S;
                gather_kernel<<<(8704 + 127) / 128, 128, 0, st[s]>>>(reinterpret_cast<const int4*>(hin[s]), reinterpret_cast<int4*>(din[s]), 8704);
                spin_kernel<<<1024, 64, 0, st[s]>>>(din[s], hout[s], ns);
            }
            t1 = now(); cudaDeviceSynchronize(); t2 = now();
            snprintf(nm, sizeof nm, "gather kernel (host->device) -> kernel(%lld ns) writing host, 8 streams", ns);
            if (rep) report(nm, t0, t1, t2, K);
            t0 = now();
            for (int k = 0; k < K; ++k) {
                const int s = k % S;
                spin_kernel<<<1024, 64, 0, st[s]>>>(hin[s], hout[s], ns);
            }
            t1 = now(); cudaDeviceSynchronize(); t2 = now();
            snprintf(nm, sizeof nm, "kernel(%lld ns) reading and writing pinned host memory, 8 streams", ns);
            if (rep) report(nm, t0, t1, t2, K);
        }
    }
    // synchronous latency of one step on one stream
    for (int mode = 0; mode < 4; ++mode) {
        const long long ns = 8000;
        double t0 = now();
        const int R = 500;
        for (int k = 0; k < R; ++k) {
            const double* b = reinterpret_cast<const double*>(hin[0]);
            if (mode == 0) { cudaMemcpyAsync(din[0], hin[0], IN, cudaMemcpyHostToDevice, st[0]); spin_kernel<<<1024, 64, 0, st[0]>>>(din[0], dout[0], ns); cudaMemcpyAsync(hout[0], dout[0], OUT, cudaMemcpyDeviceToHost, st[0]); }
            if (mode == 1) { cudaMemcpyAsync(din[0], hin[0], IN, cudaMemcpyHostToDevice, st[0]); spin_kernel<<<1024, 64, 0, st[0]>>>(din[0], hout[0], ns); }
            if (mode == 2) { gather_kernel<<<68, 128, 0, st[0]>>>(reinterpret_cast<const int4*>(hin[0]), reinterpret_cast<int4*>(din[0]), 8704); spin_kernel<<<1024, 64, 0, st[0]>>>(din[0], hout[0], ns); }
            if (mode == 3) spin3_kernel<<<1024, 64, 0, st[0]>>>(b, b + 9 * 1024, b + 12 * 1024, hout[0], ns);
            cudaStreamSynchronize(st[0]);
        }
        double t2 = now();
        const char* nm[] = {"sync: H2D -> kernel(8 us) -> D2H", "sync: H2D -> kernel(8 us) writing host", "sync: gather kernel -> kernel(8 us) writing host", "sync: kernel(8 us) reading 3 host arrays, writing host"};
        printf("%-64s %7.2f us per step\n", nm[mode], (t2 - t0) / R * 1e6);
    }
    return 0;
}
