// fs_probe.c -- what bounds the FASTQ file write on a box (measurement tool, not part of the product).
// Writes `total` bytes from a host buffer into files of a directory in several ways and prints one JSON line per case:
//   own_files   T threads, pwrite, one file per thread                (per-file locks out of the picture)
//   one_file    T threads, pwrite, ONE file, interleaved 4 MB chunks   (what a FASTQ file needs)
//   rewrite     the same over the file just written                    (pages exist: copy cost without allocation)
//   falloc      posix_fallocate of the whole file first, then one_file (allocation moved out of the writers)
//   mmap        ftruncate + mmap(MAP_SHARED) of ONE file, T threads memcpy their chunks
//   direct      O_DIRECT pwrite into ONE file (skipped where the file system refuses the flag)
// build: gcc -O2 -o tools/build/fs_probe tools/fs_probe.c -lpthread;  run: fs_probe <dir> [total_MB] [threads,...]
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define CHUNK (4u << 20)

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct { int fd; uint8_t* map; const uint8_t* src; size_t srcLen; size_t total; int t, T; int ownFile; int err; } Job;

static void* worker(void* p) {
	Job* j = (Job*)p;
	const size_t nChunks = j->total / CHUNK;
	for (size_t c = j->t; c < nChunks; c += (size_t)j->T) {
		const uint8_t* s = j->src + (c * CHUNK) % (j->srcLen - CHUNK + 1) / 4096 * 4096;
		const off_t off = j->ownFile ? (off_t)((c / (size_t)j->T) * CHUNK) : (off_t)(c * CHUNK);
		if (j->map) { memcpy(j->map + off, s, CHUNK); continue; }
		size_t done = 0;
		while (done < CHUNK) {
			ssize_t w = pwrite(j->fd, s + done, CHUNK - done, off + (off_t)done);
			if (w < 0) { if (errno == EINTR) continue; j->err = errno; return NULL; }
			done += (size_t)w;
		}
	}
	return NULL;
}

static double run(int T, int* fds, uint8_t* map, const uint8_t* src, size_t srcLen, size_t total, int ownFile, int* err) {
	pthread_t th[64]; Job jobs[64];
	const double t0 = now();
	for (int t = 0; t < T; t++) {
		jobs[t] = (Job){ownFile ? fds[t] : fds[0], map, src, srcLen, total, t, T, ownFile, 0};
		pthread_create(&th[t], NULL, worker, &jobs[t]);
	}
	for (int t = 0; t < T; t++) { pthread_join(th[t], NULL); if (jobs[t].err) *err = jobs[t].err; }
	return now() - t0;
}

static void report(const char* dir, const char* mode, int T, size_t total, double dt, int err) {
	if (err) printf("{\"dir\": \"%s\", \"mode\": \"%s\", \"threads\": %d, \"error\": \"%s\"}\n", dir, mode, T, strerror(err));
	else printf("{\"dir\": \"%s\", \"mode\": \"%s\", \"threads\": %d, \"GBps\": %.2f}\n", dir, mode, T, total / dt / 1e9);
	fflush(stdout);
}

int main(int argc, char** argv) {
	if (argc < 2) { fprintf(stderr, "usage: fs_probe <dir> [total_MB] [threads,...]\n"); return 2; }
	const char* dir = argv[1];
	const size_t total = (size_t)(argc > 2 ? atol(argv[2]) : 4096) << 20;
	int Ts[16], nT = 0;
	{ char* s = strdup(argc > 3 ? argv[3] : "1,4,16"); for (char* p = strtok(s, ","); p && nT < 16; p = strtok(NULL, ",")) Ts[nT++] = atoi(p); }
	const size_t srcLen = 256u << 20;
	uint8_t* src = NULL;
	if (posix_memalign((void**)&src, 4096, srcLen)) return 1;
	for (size_t i = 0; i < srcLen; i++) src[i] = (uint8_t)(i * 2654435761u >> 24);
	char path[64][512];
	for (int i = 0; i < 64; i++) snprintf(path[i], sizeof path[i], "%s/fsprobe_%d_%d", dir, (int)getpid(), i);
	for (int ti = 0; ti < nT; ti++) {
		const int T = Ts[ti] > 64 ? 64 : Ts[ti];
		int fds[64], err = 0; double dt;
		// own_files
		for (int t = 0; t < T; t++) fds[t] = open(path[t], O_CREAT | O_TRUNC | O_WRONLY, 0644);
		dt = run(T, fds, NULL, src, srcLen, total, 1, &err);
		for (int t = 0; t < T; t++) { close(fds[t]); unlink(path[t]); }
		report(dir, "own_files", T, total, dt, err);
		// one_file, then rewrite
		err = 0; fds[0] = open(path[0], O_CREAT | O_TRUNC | O_WRONLY, 0644);
		dt = run(T, fds, NULL, src, srcLen, total, 0, &err); report(dir, "one_file", T, total, dt, err);
		err = 0; dt = run(T, fds, NULL, src, srcLen, total, 0, &err); report(dir, "rewrite", T, total, dt, err);
		close(fds[0]); unlink(path[0]);
		// falloc
		err = 0; fds[0] = open(path[0], O_CREAT | O_TRUNC | O_WRONLY, 0644);
		{ const double t0 = now(); const int e = posix_fallocate(fds[0], 0, (off_t)total); const double ta = now() - t0;
		  dt = run(T, fds, NULL, src, srcLen, total, 0, &err);
		  if (e) err = e;
		  report(dir, "falloc(writers only)", T, total, dt, err); report(dir, "falloc(incl. fallocate)", T, total, dt + ta, err); }
		close(fds[0]); unlink(path[0]);
		// mmap
		err = 0; fds[0] = open(path[0], O_CREAT | O_TRUNC | O_RDWR, 0644);
		if (ftruncate(fds[0], (off_t)total) == 0) {
			const double t0 = now();
			uint8_t* m = (uint8_t*)mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED, fds[0], 0);
			if (m != MAP_FAILED) { run(T, fds, m, src, srcLen, total, 0, &err); munmap(m, total); dt = now() - t0; report(dir, "mmap", T, total, dt, err); }
			else report(dir, "mmap", T, total, 1, errno);
		}
		close(fds[0]); unlink(path[0]);
		// direct
		err = 0; fds[0] = open(path[0], O_CREAT | O_TRUNC | O_WRONLY | O_DIRECT, 0644);
		if (fds[0] < 0) report(dir, "direct", T, total, 1, errno);
		else { dt = run(T, fds, NULL, src, srcLen, total, 0, &err); report(dir, "direct", T, total, dt, err); close(fds[0]); }
		unlink(path[0]);
	}
	return 0;
}
