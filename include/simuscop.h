/*
 * simuscop.h -- C ABI of the B200 read-generation hot path (libsimuscop_cuda.so).
 *
 * This is the drop-in boundary for the per-read loop of SimuSCoP's `simuReads`.
 * The reference has no FFI for this path; it enters it through a C function
 * pointer and process-wide singletons:
 *
 *   threadPool->pool_add_work(&Segment::yieldReads, &chrSegs[k], n++)
 *                                     reference lib/genome/Genome.cpp:881, :949
 *   void* Segment::yieldReads(const void* seg)      lib/segment/Segment.cpp:673-871
 *   char* Profile::predict(char* refSeq, int isRead1)  lib/profile/Profile.cpp:1586-1701
 *   int   Profile::yieldInsertSize()                lib/profile/Profile.cpp:1486-1493
 *   swp->write(char*[, char*])                      lib/seqwriter/SeqWriter.cpp:41-54
 *
 * The entry points below are what a maintainer binds where Genome::yieldReads
 * dispatched segments to the thread pool (INTEGRATION.md shows the stub).
 * Conventions: plain pointers and sizes only; every function returns SSC_OK (0)
 * or an SSC_ERR_* code and records a message readable with ssc_last_error();
 * no exception crosses the boundary; host pointers are borrowed for the duration
 * of the call only; one handle per GPU; a handle is not thread-safe, distinct
 * handles may be used from distinct host threads.  There is no CPU fallback:
 * without a CUDA device ssc_create() fails.
 */
#ifndef SIMUSCOP_H
#define SIMUSCOP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSC_OK            0
#define SSC_ERR_INVALID   1   /* bad argument / unsupported configuration */
#define SSC_ERR_CUDA      2   /* CUDA runtime error */
#define SSC_ERR_STATE     3   /* call order violated (e.g. generate before set_plan) */
#define SSC_ERR_NOMEM     4
#define SSC_ERR_SINK      5   /* the sink callback returned non-zero */
#define SSC_ERR_OVERFLOW  6   /* a read outgrew the per-read scratch (limits) */

typedef struct ssc_handle ssc_handle;

/*
 * Profile tables -- the FP64 CDFs exactly as the reference builds them
 * (Profile::load -> normParas(true) -> initCDFs, lib/profile/Profile.cpp:934-1434;
 * Matrix::normalize/cumsum, lib/matrix/Matrix.h:482-522).  Row-major:
 *   subs_cdf1/2 [n_kmer_rows][bins][n_bases]      (Profile::subsCdf1/2)
 *   quality_cdf [n_bases*n_bases][bins][n_qual]   (Profile::qualityCdf)
 *   isize_cdf   [n_isize]  value k <-> insert size min_insert_size+k (Profile::iSizeCdf)
 *   ins_cdf/del_cdf [n_ins]/[n_del]               (Profile::insCdf/delCdf)
 * n_isize == 0 means "no insert-size table": every fragment uses fixed_insert_size
 * (Profile::yieldInsertSize, Profile.cpp:1487-1489).  use_cdf2 == 0 means read 2
 * uses subs_cdf1 (Profile::getSubBaseIndx2, Profile.cpp:1546-1549).
 */
typedef struct ssc_profile_tables {
	int32_t n_bases;            /* N, must be 4 */
	int32_t kmer;               /* K */
	int32_t bins;               /* B */
	int32_t n_qual;             /* Q = 94 */
	int32_t min_qual;           /* 33 */
	int32_t read_length;        /* RL */
	int32_t paired;             /* layout == "PE" */
	int32_t use_cdf2;
	int32_t fixed_insert_size;  /* config insertSize */
	int32_t min_insert_size;
	int32_t n_isize;
	int32_t n_ins;
	int32_t n_del;
	int32_t n_kmer_rows;        /* sum_{p=1..K} N^p */
	double insert_rate;
	double del_rate;
	char bases[8];              /* e.g. "ACTG", NUL padded */
	const double* isize_cdf;
	const double* ins_cdf;
	const double* del_cdf;
	const double* subs_cdf1;
	const double* subs_cdf2;
	const double* quality_cdf;
} ssc_profile_tables;

/*
 * One sampling bin of the read plan (Segment::fragStartPos/fragEndPos/hapIndxs/fragRCs,
 * lib/segment/Segment.h:44-48), flattened so that the device needs no segment objects.
 * Bins are given in the reference's threads=1 emission order
 * (population -> chromosome -> segment -> bin); bins of one segment are contiguous.
 */
typedef struct ssc_bin {
	int64_t hap_base;     /* index in the haplotype store of the first base of this
	                         segment's haplotype string (segSequences[hap]) */
	int64_t contig_end;   /* one past the last base of the (population, chromosome,
	                         haplotype index) contig: the concatenation of the same-index
	                         haplotype strings of the chromosome's segments, which is what
	                         Segment::getFragSequence + Genome::produceFragment walk
	                         (Segment.cpp:1077-1103, Genome.cpp:599-632) */
	int32_t spos, epos;   /* start position is drawn in [spos, epos] */
	uint32_t segsize;     /* seqSize/CN, the modulus of the position in read names */
	int32_t read_count;   /* fragRCs[i] */
	int32_t segment;      /* index into the segment array */
	int32_t reserved;
} ssc_bin;

typedef struct ssc_segment {
	int64_t first_bin;
	int64_t n_bins;
	int32_t name_offset;  /* "@<popu>#<chr>#" inside the name blob (Segment.cpp:780,809,824) */
	int32_t name_len;
} ssc_segment;

typedef struct ssc_stats {
	double   device_ms;        /* sum of CUDA-event durations of the generation kernels */
	uint64_t launches;         /* kernels launched by this handle since creation */
	uint64_t gen_launches;     /* launches of the fused generation kernel */
	uint64_t pairs_emitted;
	uint64_t reads_emitted;
	uint64_t bases_emitted;    /* sum of emitted read lengths */
	uint64_t fastq_bytes;      /* bytes of FASTQ written to HBM */
	uint64_t hap_bytes;        /* algorithmic haplotype bytes read: ceil(len/4)+ceil(len/8) per fragment */
	uint64_t d2h_bytes;
	uint64_t h2d_bytes;
	double   gen_kernel_ms;    /* CUDA-event time of the generation kernel alone (ssc_generate_device) */
	double   compact_kernel_ms;/* CUDA-event time of the compaction kernel alone (ssc_generate_device) */
	uint64_t timed_batches;    /* launches covered by gen_kernel_ms */
	uint64_t gz_bytes;         /* gzip mode: compressed bytes written to the slabs (fastq_bytes stays the plain size) */
	uint64_t bin_bytes;        /* algorithmic bin-record bytes read: 64 bytes per distinct bin of every generated batch */
} ssc_stats;

/*
 * Sink for finished FASTQ slabs, called in pair order.  buf1/buf2 point to pinned host
 * memory owned by the handle and valid only during the call (SeqWriter::write(char*,char*)
 * semantics, SeqWriter.cpp:49-54: the callee copies or writes synchronously).  In SE
 * layout buf2 == NULL and len2 == 0.  Return 0 to continue.
 */
typedef int (*ssc_sink_fn)(void* user, const char* buf1, size_t len1, const char* buf2, size_t len2,
                           int64_t first_pair, int64_t n_pairs);

const char* ssc_last_error(void);
int ssc_version(void);

int ssc_create(int device, ssc_handle** out);
int ssc_destroy(ssc_handle* h);

/* "batch_pairs" (pairs per kernel launch), "fp64_search" (0/1: use the FP64 linear-search
 * ground-truth kernel instead of the integer-threshold kernel), "force_generic" (0/1), "gzip" (0/1: the slabs
 * handed to the sink hold concatenated gzip members -- a valid .gz stream of the same FASTQ bytes -- compressed on
 * the GPU; the plain bytes of SeqWriter::write are the default), "carry_pass2" (1: the blobs of batch k are moved into the dense slab by the
 * generation kernel of batch k+1; default 0: by a stand-alone kernel after every batch), "concurrent_move" (n > 0: the blobs of batch k are moved by a
 * bulk-copy kernel on n SMs, on a second stream, while the generation kernel of batch k+1 runs on the others; default 8 of
 * 148 SMs; 0: a stand-alone kernel after every batch), "prefetch_windows" (default 1: the
 * ticket prologue of the generation kernel pulls the haplotype windows of its pairs into the L2), "max_ctas" (> 0 caps the grid of the generation
 * kernel; 0 = one CTA per SM), "no_splice" (0/1, tests: reads with indel events take the position-by-position path
 * of the fast kernel instead of the spliced packed read). */
int ssc_set_option(ssc_handle* h, const char* key, int64_t value);

int ssc_set_profile(ssc_handle* h, const ssc_profile_tables* t);

/* Haplotype store: 2-bit codes (in the profile's `bases` order) + 1-bit non-ACGT mask,
 * packed on the device from ASCII (any case; every non-ACGT character behaves as N).
 * Append the haplotype strings contig by contig; *first_base receives the store index of
 * the first appended base. */
int ssc_genome_reserve(ssc_handle* h, uint64_t total_bases);
int ssc_genome_append(ssc_handle* h, const char* ascii, uint64_t n, uint64_t* first_base);
int ssc_genome_size(ssc_handle* h, uint64_t* n_bases);

/* Reference-relative construction of haplotype strings on the device (the copy-and-substitute part of
 * Segment::generateSegSequences, lib/segment/Segment.cpp:124-311: a haplotype of a segment without indel variants is
 * the reference slice repeated once per copy, with SNP / SNV alleles poked in).  The chromosome is uploaded once
 * (ssc_reference_upload, any case, ASCII); ssc_genome_append_ref appends `reps` copies of reference bases
 * [ref_off, ref_off+len) to the store exactly as ssc_genome_append would append that string; ssc_genome_poke then
 * overwrites single bases (store indices, distinct) with the given characters.  ssc_genome_read decodes store bases
 * back to upper-case ASCII (non-ACGT -> 'N'), a diagnostic for tests. */
int ssc_reference_upload(ssc_handle* h, const char* ascii, uint64_t n);
/* The same from the FASTA file itself (FastaReference::getSequence, lib/fastahack/Fasta.cpp:304-334, and the upper-casing
 * of Genome::getSubSequence done on the device): the sequence lines of one record are read from `fd` (raw_len bytes at
 * file_offset, straight into pinned staging buffers) and unfolded on the GPU with the geometry of the record's .fai entry
 * (n_bases bases, line_bases bases per line in line_width bytes).  *n_other receives the number of characters that are
 * neither ACGT nor N in either case (IUPAC codes). */
int ssc_reference_upload_fasta(ssc_handle* h, int fd, uint64_t file_offset, uint64_t raw_len, uint64_t n_bases,
                               uint32_t line_bases, uint32_t line_width, uint64_t* n_other);
/* The same in the background: ssc_reference_prefetch_fasta starts reading and unfolding a record into a second set of
 * buffers (own thread, own stream) and returns; the caller goes on working with the current reference (ssc_genome_append_ref,
 * ssc_gc_census, host work) and later calls ssc_reference_adopt_prefetched, which waits for the record and makes it the
 * current reference.  One record may be in flight. */
int ssc_reference_prefetch_fasta(ssc_handle* h, int fd, uint64_t file_offset, uint64_t raw_len, uint64_t n_bases,
                                 uint32_t line_bases, uint32_t line_width);
int ssc_reference_adopt_prefetched(ssc_handle* h, uint64_t* n_other);
int ssc_genome_append_ref(ssc_handle* h, uint64_t ref_off, uint64_t len, int32_t reps, uint64_t* first_base);
int ssc_genome_poke(ssc_handle* h, const int64_t* store_pos, const char* chars, int64_t n);
int ssc_genome_read(ssc_handle* h, uint64_t start, uint64_t n, char* out);

/* GC census of haplotype-store intervals: the device half of the GC-weighted read plan
 * (Segment::getWeightedLength, lib/segment/Segment.cpp:567-624, which calls calculateGCPercent,
 * lib/mydefine/MyDefine.cpp:279-303, once per 1 kb window / capture target).  For interval i =
 * store bases [starts[i], starts[i]+lens[i]) (lens[i] >= 0, inside the appended store) it returns
 * gc[i] = number of G or C bases and nn[i] = number of non-ACGT bases; the caller forms the
 * reference's integer percentage (nn > 0 ? -1 : 100*gc/len).  Host arrays, copied inside. */
int ssc_gc_census(ssc_handle* h, const int64_t* starts, const int32_t* lens, int64_t n,
                  int32_t* gc, int32_t* nn);

/* Uploads the plan, runs the fragment census (failCount > 1000 rule, Segment.cpp:753-762)
 * for `seed`, and computes pair / fragCount prefix sums.  *planned_pairs = number of pair
 * IDs (PE: sum ceil(read_count/2); SE: sum read_count). */
int ssc_set_plan(ssc_handle* h, uint64_t seed,
                 const ssc_bin* bins, int64_t n_bins,
                 const ssc_segment* segs, int64_t n_segs,
                 const char* names, int64_t names_len,
                 int64_t* planned_pairs, int64_t* emitted_pairs);

/* Generates planned pair IDs [pair_lo, pair_hi) and streams the FASTQ bytes to `sink`
 * through pinned double buffers (host copies inside). */
int ssc_generate(ssc_handle* h, int64_t pair_lo, int64_t pair_hi, ssc_sink_fn sink, void* user);

/* Same work, output left in the handle's device slab (no device->host copy of FASTQ):
 * the device-resident measurement leg.  Any of the out pointers may be NULL. */
int ssc_generate_device(ssc_handle* h, int64_t pair_lo, int64_t pair_hi,
                        uint64_t* bytes1, uint64_t* bytes2, uint64_t* bases, double* device_ms);

/* Host-only diagnostic (no GPU needed): the symbol the integer-threshold table built from
 * `cdf` selects for the 32-bit draw u.  Must equal the reference's
 * randIndx(cdf, n) with r = 2.2204e-16 + (1 - 2.2204e-16) * (u / 2^32) for every u
 * (lib/mydefine/MyDefine.cpp:176-184); tests sweep the table boundaries with it. */
int ssc_table_lookup_host(const double* cdf, int n, uint32_t u);
/* Same for a 4-symbol substitution row encoded as {S0,S1,S2,base}. */
int ssc_sub_lookup_host(const double* cdf4, uint32_t u);

/* Host-only diagnostic (no GPU needed): one gzip member of in[0..n) built with the device encoder's own tables
 * (Huffman code fitted to the byte histogram of `sample`, CRC-32 by the kernel's lane-strided algebra) and bit
 * packing.  Returns the member size, or -1.  Tests inflate it with zlib. */
int64_t ssc_gzip_member_host(const uint8_t* in, uint32_t n, const uint8_t* sample, size_t sample_n, uint8_t* out, size_t cap);

/* Measurement aid (no reference counterpart): the instruction-issue ceiling of the generation kernel on this GPU.
 * mode 0 = Philox4x32-10 only, one block per lane-cycle in the generation kernel's launch shape (the cost of the reference's
 * four uniform draws per base, Profile.cpp:1560/1569/1534/1578, and nothing else); mode 1 = the same plus the per-base table
 * work and byte stores of an indel-free read on synthetic tables.  Runs `reps` launches over n_pairs pairs of read_length
 * cycles per mate and returns the mean CUDA-event time of a launch.  bench.py reports it as roofline.issue. */
int ssc_issue_floor(ssc_handle* h, int mode, int read_length, int64_t n_pairs, int reps, double* ms_per_launch);

int ssc_get_stats(ssc_handle* h, ssc_stats* out);
int ssc_reset_stats(ssc_handle* h);

#ifdef __cplusplus
}
#endif
#endif
