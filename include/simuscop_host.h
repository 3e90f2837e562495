/*
 * simuscop_host.h -- C ABI of the host front end (libsimuscop_host.so).
 *
 * Replaces, for the simuReads path, what the reference does between reading the
 * configuration file and dispatching segments:
 *   config.loadConfig(); genome.loadData(); profile.train(file);
 *   genome.generateSegments(); genome.yieldReads()            src/simuReads.cpp:48-74
 * ssh_open() covers everything up to generateSegments(); ssh_prepare_sample() is the
 * host half of Genome::yieldReads (read budget, GC-weighted bins, haplotype strings;
 * lib/genome/Genome.cpp:827-960) and hands the result to a device handle of
 * simuscop.h.  Samples must be prepared in increasing order (plan RNG state is shared,
 * as in the reference).  Errors in the input files terminate the process with the
 * reference's own messages and exit codes (exit(1) / exit(-1)), like the reference.
 */
#ifndef SIMUSCOP_HOST_H
#define SIMUSCOP_HOST_H

#include <stdint.h>
#include "simuscop.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssh_job ssh_job;

int ssh_open(const char* config_path, uint64_t seed, ssh_job** out);
int ssh_close(ssh_job* job);
int ssh_num_samples(ssh_job* job);
/* file-name stem of sample s: "<popu>" or "<p1>_0.300+<p2>_0.250..." (Genome.cpp:857-866, 899-929) */
const char* ssh_sample_stem(ssh_job* job, int s);
int ssh_paired(ssh_job* job);
int ssh_read_length(ssh_job* job);
const char* ssh_output_dir(ssh_job* job);
/* dump_path may be NULL; otherwise an SSCPLAN1 file of the sample is written there */
int ssh_prepare_sample(ssh_job* job, int s, ssc_handle* dev, const char* dump_path,
                       int64_t* planned_pairs, int64_t* emitted_pairs);
/* Host-only self test (no GPU): the splice lists from which the device assembles haplotypes with insertion / deletion
 * variants, materialised on the host, against the string construction of Segment::generateSegSequences
 * (lib/segment/Segment.cpp:124-460).  Returns the number of differing haplotypes. */
int ssh_selftest_splices(ssh_job* job, int64_t* haplotypes_checked, int64_t* segments_with_indels);

/* Output side (the role of SeqWriter, lib/seqwriter/SeqWriter.cpp:41-54): an ordered FASTQ file writer whose sink can be
 * handed to ssc_generate with user = the writer.  path2 NULL = single-end.  threads <= 1: slabs are written by the caller's
 * thread; otherwise one pwrite() stream per file, the two files side by side (a file is a serial resource under its inode
 * lock, so the default mode never uses more than two threads).  Environment SIMUSCOP_WRITER_MODE = mmap | hybrid: `threads`
 * workers copy 8 MB chunks into shared mappings of the slab's file ranges (hybrid: beside the two streams). */
typedef struct ssh_writer ssh_writer;
int ssh_writer_open(const char* path1, const char* path2, int threads, ssh_writer** out);
ssc_sink_fn ssh_writer_sink(void);
int ssh_writer_close(ssh_writer* w, uint64_t* bytes1, uint64_t* bytes2);

/* The whole drop-in run: every sample -> FASTQ files in the output directory, on `device`. */
int ssh_run(ssh_job* job, int device);

#ifdef __cplusplus
}
#endif
#endif
