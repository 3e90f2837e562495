/*
 * ssc_hooks.h -- TEST INFRASTRUCTURE (oracle side). Not part of the product.
 *
 * Declarations of the stream-addressing hooks that oracle/build_ref.py inserts
 * into a scratch copy of the reference sources when it builds the
 * Philox-instrumented reference binary (oracle/_ref/simuReads_philox).
 * The hooks tell the replacement RNG (philox_pool.cpp) which reference draw
 * site is about to consume a uniform, so that every draw maps to a fixed
 * (pairID, mate, stream, index, word) Philox address -- DESIGN.md "Philox
 * addressing".  Nothing else of the reference is changed.
 */
#ifndef SSC_HOOKS_H
#define SSC_HOOKS_H

#include <string>
#include <vector>

enum SscSlot {
	SSC_SLOT_NONE = 0,
	SSC_SLOT_POS,      /* Segment.cpp:743  randomInteger(spos, epos+1)      -> stream 0, x0 */
	SSC_SLOT_ISIZE,    /* Profile.cpp:1491 randIndx(iSizeCdf)               -> stream 0, x1 */
	SSC_SLOT_STRAND,   /* Segment.cpp:766  randomInteger(0, 2) (SE)         -> stream 0, x2 */
	SSC_SLOT_P,        /* Profile.cpp:1560 randomDouble(0,1) insertion test -> stream 1, x0 */
	SSC_SLOT_P2,       /* Profile.cpp:1569 randomDouble(0,1) deletion test  -> stream 1, x1 */
	SSC_SLOT_SUB,      /* Profile.cpp:1534/1547/1551 randIndx(subsCdf)      -> stream 1, x2 */
	SSC_SLOT_QUAL,     /* Profile.cpp:1578 randIndx(qualityCdf) / :1583     -> stream 1, x3 */
	SSC_SLOT_INSLEN,   /* Profile.cpp:1520 randIndx(insCdf)                 -> stream 2, x0 */
	SSC_SLOT_DELLEN,   /* Profile.cpp:1524 randIndx(delCdf)                 -> stream 2, x1 */
	SSC_SLOT_INSBASE   /* Profile.cpp:1564 randomInteger(0, N-1)            -> stream 3, word i */
};

void ssc_hook_bin(int readCount);     /* top of the per-bin loop, Segment.cpp:741 */
void ssc_hook_attempt();              /* top of while(n > 0), Segment.cpp:742 */
void ssc_hook_frag_ok();              /* fragCount++, Segment.cpp:763 */
void ssc_hook_slot(int slot);         /* function-entry hooks of the samplers */
void ssc_hook_mate(int isRead1);      /* Profile::predict entry */
void ssc_hook_refpos(int j);          /* indel loop body, Profile.cpp:1611 */
void ssc_hook_outpos(int j);          /* output loop body, Profile.cpp:1671 */
unsigned ssc_hook_gc_seed(unsigned l);/* Profile.cpp:1410 */
unsigned ssc_hook_plan_seed();        /* Genome.cpp:852 */
void ssc_hook_sample_end();           /* before "delete swp", Genome.cpp:888,957 */

class Segment;
void ssc_hook_dump_chr(std::vector<Segment>& chrSegs, const std::string& popu, const std::string& chr);

/* consumed by the replacement ThreadPool::randomDouble / randomInteger */
unsigned int ssc_next_u32(bool integerDraw);

#endif
