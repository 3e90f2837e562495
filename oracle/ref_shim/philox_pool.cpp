/*
 * philox_pool.cpp -- TEST INFRASTRUCTURE (oracle side). Not part of the product.
 *
 * Replacement implementation of the reference's ThreadPool class
 * (declared in /root/reference/lib/threadpool/ThreadPool.h) used ONLY for the
 * Philox-instrumented reference binary.  Differences from the reference's
 * lib/threadpool/ThreadPool.cpp:
 *   - work items run synchronously inside pool_add_work(), so the emission
 *     order is the canonical "threads = 1" order (segment -> bin -> draw);
 *   - randomDouble()/randomInteger() take their 32-bit uniform from the
 *     addressed Philox word (ssc_next_u32) instead of a per-thread mt19937.
 * The u32 -> value map is the reference's expression, kept verbatim in shape
 * and types (ThreadPool.cpp:203-212): start+(end-start)*((u-min)/(max-min+1.0)).
 */
#include "ThreadPool.h"
#include "ssc_hooks.h"

void ThreadPool::pool_init() {
	minRandNumber = 0;
	maxRandNumber = 4294967295L;   /* mt19937::min()/max() in the reference */
	work_list = NULL;
	rear = NULL;
	cur_queue_size = 0;
	finishedWorks = 0;
	shutdown = false;
}

void ThreadPool::pool_destroy() {}

void ThreadPool::clearWorks() {}

ThreadPool::~ThreadPool() {}

void ThreadPool::thread_routine() {}

void ThreadPool::pool_add_work(void *(*process)(const void *arg), const void *arg, int wid) {
	(void)wid;
	(*process)(arg);
	finishedWorks++;
}

void ThreadPool::wait() {}

double ThreadPool::randomDouble(double start, double end) {
	double number = ssc_next_u32(false);
	return start+(end-start)*((number-minRandNumber)/(maxRandNumber-minRandNumber+1.0));
}

long ThreadPool::randomInteger(long start, long end) {
	double number = ssc_next_u32(true);
	return start+(end-start)*((number-minRandNumber)/(maxRandNumber-minRandNumber+1.0));
}
