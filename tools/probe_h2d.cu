// Host-to-device ceiling probe: bare cudaMemcpyAsync from pinned host memory on every visible GPU AT ONCE.
//
//   nvcc -O2 -std=c++17 -o tools/probe_h2d tools/probe_h2d.cu -lpthread
//   tools/probe_h2d [MiB per copy = 1024] [copies = 8] [numa = 1] [write_combined = 0]
//
// bench.py's `e2e` number is bound by these copies (55 GB/s per GPU alone; less per GPU when several copy at once on
// this pool's boxes).  This program takes the framework out of the picture: one thread per GPU, one pinned buffer and
// one stream each, all threads released together, every copy timed with CUDA events.  With numa=1 each thread first
// binds itself to the CPUs of its GPU's NUMA node (sysfs), so the pinned pages are allocated there.
// Output: one JSON line {"gpus": N, "per_gpu_GBps": [...], "aggregate_GBps": ..., "numa_nodes": [...]}.
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <cctype>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

static int numa_node_of(int dev) {
    char bus[64];
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static bool bind_to_node(int node) {
    if (node < 0) return false;
    std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096];
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    char* p = buf;
    while (*p) {
        char* end;
        long lo = strtol(p, &end, 10), hi = lo;
        if (end == p) break;
        if (*end == '-') hi = strtol(end + 1, &end, 10);
        for (long c = lo; c <= hi; ++c) CPU_SET((int)c, &set);
        p = (*end == ',') ? end + 1 : end;
        if (*end != ',') break;
    }
    return sched_setaffinity(0, sizeof(set), &set) == 0;
}

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? (size_t)atoll(argv[1]) : 1024;
    const int copies = argc > 2 ? atoi(argv[2]) : 8;
    const int numa = argc > 3 ? atoi(argv[3]) : 1;
    const int wc = argc > 4 ? atoi(argv[4]) : 0;   // cudaHostAllocWriteCombined: not snooped during the DMA reads
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    const size_t bytes = mib << 20;
    std::vector<double> gbps(n, 0.0);
    std::vector<int> nodes(n, -1);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> threads;
    for (int d = 0; d < n; ++d) {
        threads.emplace_back([&, d]() {
            cudaSetDevice(d);
            nodes[d] = numa_node_of(d);
            if (numa) bind_to_node(nodes[d]);
            void *h = nullptr, *g = nullptr;
            cudaStream_t st;
            cudaEvent_t e0, e1;
            if (cudaHostAlloc(&h, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&g, bytes) != cudaSuccess) {
                ready++;
                return;
            }
            memset(h, 1, bytes);   // first touch on this thread's node
            cudaStreamCreate(&st);
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaMemcpyAsync(g, h, bytes, cudaMemcpyHostToDevice, st);   // warm-up
            cudaStreamSynchronize(st);
            ready++;
            while (!go.load()) std::this_thread::yield();
            cudaEventRecord(e0, st);
            for (int i = 0; i < copies; ++i) cudaMemcpyAsync(g, h, bytes, cudaMemcpyHostToDevice, st);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            gbps[d] = (double)bytes * copies / (ms * 1e-3) / 1e9;
            cudaFreeHost(h);
            cudaFree(g);
        });
    }
    while (ready.load() < n) std::this_thread::yield();
    go.store(true);
    for (auto& t : threads) t.join();
    double total = 0.0;
    printf("{\"gpus\": %d, \"MiB_per_copy\": %zu, \"copies\": %d, \"numa_bind\": %d, \"write_combined\": %d, \"per_gpu_GBps\": [", n, mib, copies, numa, wc);
    for (int d = 0; d < n; ++d) {
        printf("%s%.2f", d ? ", " : "", gbps[d]);
        total += gbps[d];
    }
    printf("], \"aggregate_GBps\": %.2f, \"numa_nodes\": [", total);
    for (int d = 0; d < n; ++d) printf("%s%d", d ? ", " : "", nodes[d]);
    printf("]}\n");
    return 0;
}
