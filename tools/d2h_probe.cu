// d2h_probe.cu -- pure-copy ceiling of the box: N GPUs copying device slabs into pinned host memory at the same time,
// no kernels.  What the end-to-end leg of bench.py can reach at most (every FASTQ byte crosses PCIe into host memory).
//
//   nvcc -O2 -o tools/build/d2h_probe tools/d2h_probe.cu -lpthread
//   tools/build/d2h_probe [--gpus 1,2,4,8] [--bytes 1073741824] [--seconds 1.5]
//
// For every GPU count G it runs, with one worker per GPU started behind a common barrier:
//   workers   = processes (fork before any CUDA call: bench.py's one-process-per-GPU shape) | threads of one process (the CLI's)
//   alloc     = cudaHostAllocDefault | cudaHostAllocPortable | cudaHostAllocWriteCombined
//   streams   = 1 | 2 per GPU (two halves of the slab in flight at once)
//   zero-copy = a kernel storing straight into the mapped host slab (what a fused move-to-host would do) instead of the DMA
// and prints one JSON line per combination: aggregate GB/s = bytes all workers moved / (latest end - earliest start).
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Shared {
	pthread_barrier_t bar;
	double t0[16], t1[16];
	double bytes[16];
	int err[16];
};

struct Cfg { int G; bool procs; unsigned flags; const char* allocName; int streams; bool zero; size_t bytes; double seconds; };

__global__ void store_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static void worker(int g, const Cfg& c, Shared* sh) {
	sh->err[g] = 0;
	uint8_t *d = nullptr, *h = nullptr, *hd = nullptr;
	cudaStream_t st[2];
	auto ck = [&](cudaError_t e, int line) { if (e != cudaSuccess && !sh->err[g]) { sh->err[g] = line; fprintf(stderr, "gpu %d line %d: %s\n", g, line, cudaGetErrorString(e)); } };
	ck(cudaSetDevice(g), __LINE__);
	ck(cudaMalloc((void**)&d, c.bytes), __LINE__);
	ck(cudaMemset(d, 0x41, c.bytes), __LINE__);
	ck(cudaHostAlloc((void**)&h, c.bytes, c.flags | (c.zero ? cudaHostAllocMapped : 0)), __LINE__);
	if (h && !(c.flags & cudaHostAllocWriteCombined)) memset(h, 1, c.bytes);      // first touch by this worker
	if (c.zero) ck(cudaHostGetDevicePointer((void**)&hd, h, 0), __LINE__);
	for (int s = 0; s < c.streams; s++) ck(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking), __LINE__);
	const size_t part = c.bytes / c.streams / 16 * 16;
	auto one_round = [&]() {
		for (int s = 0; s < c.streams; s++) {
			if (c.zero) store_kernel<<<296, 256, 0, st[s]>>>((const uint4*)(d + s * part), (uint4*)(hd + s * part), part / 16);
			else ck(cudaMemcpyAsync(h + s * part, d + s * part, part, cudaMemcpyDeviceToHost, st[s]), __LINE__);
		}
	};
	one_round();
	ck(cudaDeviceSynchronize(), __LINE__);
	pthread_barrier_wait(&sh->bar);
	const double t0 = now_s();
	double moved = 0;
	while (now_s() - t0 < c.seconds && !sh->err[g]) {
		one_round();
		for (int s = 0; s < c.streams; s++) ck(cudaStreamSynchronize(st[s]), __LINE__);
		moved += (double)part * c.streams;
	}
	sh->t0[g] = t0; sh->t1[g] = now_s(); sh->bytes[g] = moved;
	pthread_barrier_wait(&sh->bar);
	for (int s = 0; s < c.streams; s++) cudaStreamDestroy(st[s]);
	cudaFreeHost(h); cudaFree(d);
}

struct ThreadArg { int g; const Cfg* c; Shared* sh; };
static void* thread_main(void* p) { ThreadArg* a = (ThreadArg*)p; worker(a->g, *a->c, a->sh); return nullptr; }

static void run(const Cfg& c, Shared* sh) {
	pthread_barrierattr_t ba;
	pthread_barrierattr_init(&ba);
	pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
	pthread_barrier_init(&sh->bar, &ba, c.G);
	if (c.procs) {
		std::vector<pid_t> kids;
		for (int g = 0; g < c.G; g++) {
			pid_t p = fork();
			if (p == 0) { worker(g, c, sh); _exit(0); }
			kids.push_back(p);
		}
		for (pid_t p : kids) { int st; waitpid(p, &st, 0); }
	} else {
		// threads need a CUDA-free parent too (later process rounds fork): run them in one forked child
		pid_t p = fork();
		if (p == 0) {
			std::vector<pthread_t> th(c.G);
			std::vector<ThreadArg> args(c.G);
			for (int g = 0; g < c.G; g++) { args[g] = ThreadArg{g, &c, sh}; pthread_create(&th[g], nullptr, thread_main, &args[g]); }
			for (int g = 0; g < c.G; g++) pthread_join(th[g], nullptr);
			_exit(0);
		}
		int st; waitpid(p, &st, 0);
	}
	pthread_barrier_destroy(&sh->bar);
	double t0 = 1e300, t1 = 0, bytes = 0, minG = 1e300, maxG = 0; int err = 0;
	for (int g = 0; g < c.G; g++) {
		t0 = sh->t0[g] < t0 ? sh->t0[g] : t0; t1 = sh->t1[g] > t1 ? sh->t1[g] : t1; bytes += sh->bytes[g]; err |= sh->err[g];
		const double r = sh->bytes[g] / (sh->t1[g] - sh->t0[g]) / 1e9;
		minG = r < minG ? r : minG; maxG = r > maxG ? r : maxG;
	}
	printf("{\"gpus\": %d, \"workers\": \"%s\", \"alloc\": \"%s\", \"streams\": %d, \"path\": \"%s\", \"aggregate_GBps\": %.2f, "
	       "\"per_gpu_min_GBps\": %.2f, \"per_gpu_max_GBps\": %.2f, \"error\": %d}\n",
	       c.G, c.procs ? "processes" : "threads", c.allocName, c.streams, c.zero ? "kernel stores into mapped host memory" : "cudaMemcpyAsync",
	       bytes / (t1 - t0) / 1e9, minG, maxG, err);
	fflush(stdout);
}

int main(int argc, char** argv) {
	std::string gpus = "1";
	size_t bytes = 1ull << 30;
	double seconds = 1.5;
	bool quick = false;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = argv[++i];
		else if (!strcmp(argv[i], "--bytes") && i + 1 < argc) bytes = strtoull(argv[++i], nullptr, 10);
		else if (!strcmp(argv[i], "--seconds") && i + 1 < argc) seconds = atof(argv[++i]);
		else if (!strcmp(argv[i], "--quick")) quick = true;
	}
	Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
	std::vector<int> Gs;
	for (size_t p = 0; p < gpus.size();) { size_t q = gpus.find(',', p); if (q == std::string::npos) q = gpus.size(); Gs.push_back(atoi(gpus.substr(p, q - p).c_str())); p = q + 1; }
	struct A { unsigned f; const char* n; };
	const A allocs[3] = {{cudaHostAllocDefault, "default"}, {cudaHostAllocPortable, "portable"}, {cudaHostAllocWriteCombined, "write-combined"}};
	for (int G : Gs) {
		if (G < 1 || G > 16) continue;
		for (int procs = 1; procs >= 0; procs--)
			for (int a = 0; a < 3; a++)
				for (int streams = 1; streams <= 2; streams++) {
					if (quick && (a == 1 || (procs == 0 && a != 0))) continue;
					run(Cfg{G, procs != 0, allocs[a].f, allocs[a].n, streams, false, bytes, seconds}, sh);
				}
		run(Cfg{G, true, cudaHostAllocDefault, "default", 2, true, bytes, seconds}, sh);
	}
	return 0;
}
