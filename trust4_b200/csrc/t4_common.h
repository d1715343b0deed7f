// Shared definitions of the stream engine: device-resident data layout and the
// build-mode macros.  The engine source (t4_engine.h) is written once against a
// (tid, nt, barrier) abstraction:
//   * product build (nvcc, sm_100a): one CTA per stream, T4_SYNC = __syncthreads();
//   * test-only emulation build (g++, -DT4_EMU): nt = 1, barriers are no-ops.  It
//     exists so the bit-exact logic can be debugged in a container without a GPU
//     (tests/emu); it is never linked into libtrust4_b200.so.
#ifndef T4_COMMON_H
#define T4_COMMON_H

#include <stdint.h>
#include <string.h>
#include "../../include/trust4_b200.h"

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

#if defined(__CUDACC__) && !defined(T4_EMU)
#define T4_CUDA 1
#define T4_HD __host__ __device__
#define T4_D __device__
#define T4_SYNC() __syncthreads()
#else
#define T4_CUDA 0
#define T4_HD
#define T4_D
#define T4_SYNC() ((void)0)
#endif

#define T4_DEV_MAX_READ T4_MAX_READ_LEN       /* device-side read length limit (reads > 200 bp make the reference switch to isLongSeqSet) */
#define T4_ALIGN 16
#define T4_BIG_REPEAT 10000       /* SeqSet.hpp:799, 875, 937: hits[k].repeats <= 10000 */

// ---- key layout of a seed hit (one u64 per hit; SeqSet.hpp:53 `_hit` carries the same information) ----
// [63] strand (+1 -> 1, -1 -> 0) | [62:41] contig slot | [40:20] diagonal c = a - b + 2^20 | [19:1] contig offset b | [0] repeats > 10000
#define T4_KEY_STRAND_SHIFT 63
#define T4_KEY_IDX_SHIFT 41
#define T4_KEY_C_SHIFT 20
#define T4_KEY_B_SHIFT 1
#define T4_KEY_IDX_BITS 22
#define T4_KEY_C_BIAS (1 << 20)
#define T4_KEY_B_MASK ((1u << 19) - 1)
#define T4_KEY_C_MASK ((1u << 21) - 1)
#define T4_KEY_INVALID (~0ull)

struct T4Global            // arena header at offset 0
{
	u64 top ;              // bump pointer (bytes), device-side atomicAdd
	u64 cap ;
	u64 counters[T4_N_COUNTERS] ;
	u64 firstError ;       // first T4_E_* raised by any stream since the last reset: (u32)code | aux << 32; 0 = none
	u64 pad[5] ;
} ;

struct T4Dir               // one k-mer directory slot (32 B = one HBM sector); KmerIndex.hpp:20-116
{
	u64 key ;              // k-mer code (+ barcode salt) + 1; 0 = empty
	u64 listOff ;          // arena offset of the postings array (u64 each: idx<<32 | offset)
	u32 cnt ;
	u32 cap ;
	u32 lock ;
	u32 pad ;
} ;

struct T4Contig            // SeqSet.hpp:19 `_seqWrapper`, novel contigs only (isRef is always false in the stage-1 seqSet)
{
	u64 consOff ;          // char[cap], the consensus occupies [lead, lead + len)
	u64 pwOff ;            // int32[4] per base, same lead
	u64 nameOff ;
	int nameLen ;
	int len ;              // consensusLen; alive iff consOff != 0
	int cap ;
	int lead ;             // slack on the left: left extension is a pointer move, not a memmove
	int minLeftExtAnchor, minRightExtAnchor ;
	int barcode ;
	int numRead ;
	int flags ;            // T4_CF_*
	int pad_ ;
} ;
#define T4_CF_NOINDEX 1    /* seqs[i].index == false: purged by ReleaseFinishedBarcodeSeq (SeqSet.hpp:10849) */

// bytes of one packed contig record (t4_streams_pack_contigs)
T4_HD inline u64 t4_pack_record_bytes( const T4Contig &k ) { return ( 32ull + 17ull * k.len + k.nameLen + 15 ) & ~15ull ; }

struct T4Ovl               // SeqSet.hpp:76 `_overlap`
{
	int seqIdx ;
	int readStart, readEnd ;
	int seqStart, seqEnd ;
	int strand ;
	int matchCnt ;
	int indelCnt ;
	double similarity ;
	int hcStart, hcCnt ;   // hitCoords = keys[hcStart .. hcStart+hcCnt) of the sorted hit array
	int preMatchCnt ;      // matchCnt before scoring (2*hitLen), consulted by the pre-filters SeqSet.hpp:1705-1794
	int infoFromHits ;
} ;

struct T4Stream            // one SeqSet (SeqSet.hpp:189-230 private members) + its scratch + loop state
{
	int kmerLength ;
	int radius ;
	int hitLenRequired ;
	int nomatchGapLimit ;
	int isLongSeqSet ;
	int considerBarcode ;
	double novelSeqSimilarity ;
	double repeatSimilarity ;
	// contigs
	int nSeqs, seqCap ;
	u64 seqsOff ;
	// k-mer directory
	u64 dirOff ;
	u32 dirCap, dirUsed ;
	// prevAddInfo (SeqSet.hpp:205)
	int prevSeqIdx, prevReadStart, prevReadEnd, prevSeqStart, prevStrand ;
	// scratch
	u64 keysAOff, keysBOff ;       // u64[hitCap] each
	u64 grpOff, runOff ;           // u32[hitCap + 1]
	u32 hitCap ;
	u64 keysROff, keysR2Off ;      // u64[hitCapR]: the hits once more in SortHits order, only when a k-mer has > 10000 postings
	u32 hitCapR ;
	u64 posOff ;                   // per read position scratch, see T4Pos
	u64 ovlOff, ovlTmpOff, extOff, failOff, anchorOff ;
	u64 bitsOff ;                  // u32[ovlCap * 32]: IsBaseEqual bit masks of the overhangs (ExtendOverlap)
	u32 ovlCap ;
	u64 dpOff ;                    // per-thread DP scratch, dpStride bytes each
	u32 dpStride ;
	int nThreads ;
	// driver-loop state (main.cpp:1530-1531, 640-641)
	int assembledReadCnt ;
	int prevAddRet ;
	int changeKThreshold ;
	int error ;                    // first T4_E_* raised on the device
	int errorAux ;
	// stream-local slab carved out of the arena (keeps small allocations off the global bump pointer)
	u64 slabTop, slabEnd ;
	// stats
	u64 nReads, nAddRead ;
} ;

struct T4Pos               // per (strand pass, read position) lookup record
{
	u64 listOff ;
	u64 code ;                     // rolling k-mer code ending at this position (N encoded as A)
	u32 cnt ;                      // postings of the k-mer (0 when the k-mer contains an N)
	u32 base ;                     // first hit slot, 0xffffffff = lookup not taken
} ;

// Pointers to buffers outside the arena (workloads, staging, outputs) are absolute addresses stored in u64.
template <class T> T4_HD inline T *t4_x( u64 p ) { return (T *)(uintptr_t)p ; }

struct T4Names             // gene-name table of a workload (absolute pointers)
{
	u64 pool ;                     // char[]
	u64 off ;                      // u32[n+1]
	int n ;
	int pad ;
} ;

// One op of the stream kernel.
enum
{
	T4_OP_NONE = 0,
	T4_OP_ADD_READ,
	T4_OP_REPEAT,
	T4_OP_INPUT_NOVEL,
	T4_OP_UPDATE_ALL,
	T4_OP_CHANGE_K,
	T4_OP_GET_HITS,
	T4_OP_GET_OVERLAPS,
	T4_OP_RUN_LOOP,
	T4_OP_PROBE_ONLY,
	T4_OP_INIT,
	T4_OP_RELEASE_BARCODE,
	T4_OP_RELEASE_SHALLOW,
} ;

struct T4Op                // per-CTA launch record
{
	u64 streamOff ;                // arena offset of the T4Stream
	int op ;
	int n ;                        // RUN_LOOP / PROBE_ONLY: number of descs
	u64 desc ;                     // t4_read_desc[n]            (absolute)
	u64 pool ;                     // read pool                  (absolute)
	u64 names ;                    // T4Names                    (absolute)
	u64 retCodes, strands, rescueRet ; // outputs int32[n], int8[n], int32[n] (absolute)
	u64 rescueList ;               // int32[n] scratch
	u64 good ;                     // int8[n] goodCandidate
	u64 info ;                     // int32[n]
	t4_run_cfg cfg ;
	// single-call ops
	u64 read ;                     // char[len] (absolute)
	int len ;
	int strand ;
	int barcode ;
	int minKmerCount ;
	int repetitive ;
	int kl ;
	double thr ;
	char gene[8] ;
	u64 name ; int nameLen ; int pad0 ;
	u64 out ;                      // GET_HITS / GET_OVERLAPS output buffer (absolute)
	int outCap ;
	int ret ;                      // result
	int strandOut ;
	int pad1 ;
	u64 out2 ;
	u64 stat0, stat1 ;             // PROBE_ONLY: algorithmic bytes, hits
} ;

struct T4InitParams        // T4_OP_INIT: lay out and initialise a fresh stream at streamOff
{
	int kmerLength ;
	int nomatchGapLimit ;
	int nThreads ;
	u32 seqCap, dirCap, hitCap, ovlCap ;
	u32 footprint ;
} ;

#endif
