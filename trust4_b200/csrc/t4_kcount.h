// Canonical k-mer counting and per-read count statistics on the device (the counting part of the
// stage-1 pre-processing):
//
//   KmerCount kmerCount( 21 ) ; kmerCount.AddCount( read ) for every read           KmerCount.hpp:64-97, main.cpp:404-440
//   kmerCount.GetCountStatsAndTrim( read, qual, minCnt, medianCnt, avgCnt )          KmerCount.hpp:177-288, main.cpp:981-1010
//
// (qual == NULL with --trimLevel 0; otherwise the low-quality tail of the read is cut first, KmerCount.hpp:241-271, and
// the statistics cover what is left).  minCnt / medianCnt / avgCnt order the reads
// (main.cpp:103-125) and set the similarity thresholds of the AddRead loop (main.cpp:1675-1694).
//
// The reference keeps 1 000 003 std::maps; here the counts live in one open-addressing table in HBM
// (u64 key = canonical code + 1, u32 count; load factor <= 1/2) filled with one atomicCAS + one atomicAdd per k-mer.
// Both kernels are written against the engine's (tid, nt, barrier) abstraction, so the same source runs as CTAs of
// 128 threads on the GPU and as one emulated thread in the test-only emulation build:
//   count  a CTA takes reads in turn; thread q encodes the k-mer at position q (2 bits per base, N-free window
//          = KmerCode::IsValid), takes min(code, reverse complement) = GetCanonicalKmerCode and inserts it;
//   stats  a CTA takes reads in turn; thread q looks the k-mer at position q up into shared memory, then the
//          median is found by rank counting (the element with exactly m/2 smaller-or-earlier elements is
//          c[m/2] of the sorted array), min and sum by a CTA reduction; the N rules of KmerCount.hpp:275-283 follow.
#ifndef T4_KCOUNT_H
#define T4_KCOUNT_H

#include "t4_engine.h"

struct T4KcParams
{
	u64 keys ;             // u64[cap]  (absolute device pointers), 0 = empty, else canonical code + 1
	u64 counts ;           // u32[cap]
	u64 cap ;              // power of two
	u64 pool ;             // ASCII reads
	u64 seqOff ;           // u64[n]
	u64 len ;              // i32[n]
	u64 minCnt, medianCnt ; // i32[n] out
	u64 avgCnt ;           // f32[n] out
	u64 qual ;             // ASCII qualities, same offsets as the reads; 0 = GetCountStatsAndTrim( read, NULL, ... ): no trimming
	u64 newLen ;           // i32[n] out (may be 0): length of the read after the quality trimming
	u64 ctrl ;             // u64[4]: [0] read cursor, [1] k-mers inserted, [2] distinct k-mers, [3] table full flag
	i64 n ;
	int k ;
	int pad ;
} ;

#define T4_KC_MAX_POS T4_DEV_MAX_READ
#define T4_KC_MAX_PROBES 4096      /* linear-probe bound: at load <= 1/2 clusters are a few slots long; beyond this the table is treated as full */

struct T4KcSmem
{
	char read[T4_DEV_MAX_READ + 8] ;
	int c[T4_KC_MAX_POS] ;         // counts of the valid k-mers of the current read, in position order
	u32 valid[T4_KC_MAX_POS] ;     // exclusive prefix: slot of position q among the valid ones
	u32 scan[T4_MAX_NT + 4] ;
	u64 bu[2] ;
	int bi[8] ;
} ;

T4_HD inline u64 t4_kc_hash( u64 key, u64 cap ) { return ( ( key * 0x9E3779B97F4A7C15ull ) >> 20 ) & ( cap - 1 ) ; }

// KmerCode::Append over s[q .. q + k) + GetCanonicalKmerCode (KmerCode.hpp:94-109, 52-67).  false: the window holds an N.
T4_HD inline bool t4_kc_canonical( const char *s, int q, int k, u64 *out )
{
	u64 fw = 0, rc = 0 ;
	for ( int j = 0 ; j < k ; ++j )
	{
		const char ch = s[q + j] ;
		if ( ch == 'N' )
			return false ;
		const u64 x = (u64)t4_nuc( ch ) ;
		fw = ( fw << 2 ) | x ;
		rc |= ( 3ull - x ) << ( 2 * j ) ;
	}
	*out = rc < fw ? rc : fw ;
	return true ;
}

struct T4KcCtx
{
	T4KcSmem *sm ;
	int tid, nt ;
} ;

#if T4_CUDA
#define T4_KC_SYNC() __syncthreads()
#else
#define T4_KC_SYNC() ((void)0)
#endif

// next read of this CTA (atomic cursor); -1 = none left.  Collective.
T4_D inline i64 kc_next_read( T4KcCtx &cx, const T4KcParams &P )
{
	T4_KC_SYNC() ;
	if ( cx.tid == 0 )
		cx.sm->bu[0] = t4_atomic_add( t4_x<u64>( P.ctrl ), 1ull ) ;
	T4_KC_SYNC() ;
	const u64 r = cx.sm->bu[0] ;
	return r < (u64)P.n ? (i64)r : -1 ;
}

T4_D inline void kc_load_read( T4KcCtx &cx, const T4KcParams &P, i64 r, int len )
{
	const char *src = t4_x<char>( P.pool ) + t4_x<u64>( P.seqOff )[r] ;
	for ( int i = cx.tid ; i < len ; i += cx.nt )
		cx.sm->read[i] = src[i] ;
	T4_KC_SYNC() ;
}

// KmerCount::AddCount for the reads this CTA draws
T4_D inline void kc_count_body( T4KcCtx &cx, const T4KcParams &P )
{
	u64 *keys = t4_x<u64>( P.keys ) ;
	u32 *counts = t4_x<u32>( P.counts ) ;
	u64 *ctrl = t4_x<u64>( P.ctrl ) ;
	u64 inserted = 0, fresh = 0 ;
	for ( i64 r = kc_next_read( cx, P ) ; r >= 0 ; r = kc_next_read( cx, P ) )
	{
		const int len = t4_x<int32_t>( P.len )[r] ;
		if ( len < P.k || len > T4_DEV_MAX_READ )
			continue ;
		kc_load_read( cx, P, r, len ) ;
		const int m = len - P.k + 1 ;
		for ( int q = cx.tid ; q < m ; q += cx.nt )
		{
			u64 code ;
			if ( !t4_kc_canonical( cx.sm->read, q, P.k, &code ) )
				continue ;
			const u64 key = code + 1 ;
			u64 s = t4_kc_hash( key, P.cap ) ;
			u64 probes = 0 ;
			while ( 1 )
			{
				u64 cur = keys[s] ;
				if ( cur == 0 )
				{
					cur = t4_atomic_cas( keys + s, 0ull, key ) ;
					if ( cur == 0 )
					{
						cur = key ;
						++fresh ;
					}
				}
				if ( cur == key )
				{
					t4_atomic_add32( counts + s, 1u ) ;
					++inserted ;
					break ;
				}
				s = ( s + 1 ) & ( P.cap - 1 ) ;
				if ( ++probes > T4_KC_MAX_PROBES )
				{
					ctrl[3] = 1 ; // table (nearly) full: the caller's capacity hint was too small; the results are flagged invalid
					break ;
				}
			}
		}
	}
	if ( inserted )
		t4_atomic_add( ctrl + 1, inserted ) ;
	if ( fresh )
		t4_atomic_add( ctrl + 2, fresh ) ;
}

T4_D inline u32 kc_lookup( const T4KcParams &P, u64 code )
{
	const u64 *keys = t4_x<u64>( P.keys ) ;
	const u64 key = code + 1 ;
	u64 s = t4_kc_hash( key, P.cap ) ;
	for ( int probes = 0 ; probes <= T4_KC_MAX_PROBES ; ++probes ) // bounded like the insert: an overfull table cannot hang the launch
	{
		const u64 cur = keys[s] ;
		if ( cur == key )
			return t4_x<u32>( P.counts )[s] ;
		if ( cur == 0 )
			return 0 ;
		s = ( s + 1 ) & ( P.cap - 1 ) ;
	}
	return 0 ;
}

// exclusive scan of one value per thread + total (same contract as the engine's c_scan_threads, on T4KcSmem)
T4_D inline u32 kc_scan( T4KcCtx &cx, u32 v, u32 &total )
{
	T4_KC_SYNC() ;
	cx.sm->scan[cx.tid] = v ;
	T4_KC_SYNC() ;
	u32 base = 0, tot = 0 ;
	for ( int t = 0 ; t < cx.nt ; ++t )
	{
		const u32 x = cx.sm->scan[t] ;
		if ( t < cx.tid )
			base += x ;
		tot += x ;
	}
	total = tot ;
	T4_KC_SYNC() ;
	return base ;
}

// KmerCount::GetCountStatsAndTrim( read, NULL, ... ) for the reads this CTA draws
T4_D inline void kc_stats_body( T4KcCtx &cx, const T4KcParams &P )
{
	T4KcSmem *sm = cx.sm ;
	int32_t *minCnt = t4_x<int32_t>( P.minCnt ), *medianCnt = t4_x<int32_t>( P.medianCnt ) ;
	float *avgCnt = t4_x<float>( P.avgCnt ) ;
	for ( i64 r = kc_next_read( cx, P ) ; r >= 0 ; r = kc_next_read( cx, P ) )
	{
		const int len = t4_x<int32_t>( P.len )[r] ;
		if ( len < P.k || len > T4_DEV_MAX_READ )
		{
			if ( cx.tid == 0 )
			{
				minCnt[r] = medianCnt[r] = -1 ; // KmerCount.hpp:196-200 (longer than the device limit: not supported, flagged by the host)
				avgCnt[r] = -1.0f ;
				if ( P.newLen )
					t4_x<int32_t>( P.newLen )[r] = len ;
			}
			continue ;
		}
		kc_load_read( cx, P, r, len ) ;
		const int m = len - P.k + 1 ;
		// counts of the valid k-mers, compacted in position order (the reference's c[0..k))
		// pass 1: validity + count per position (a thread owns a contiguous block of positions so the compaction is a scan)
		const int chunk = ( m + cx.nt - 1 ) / cx.nt ;
		int a = chunk * cx.tid, b = a + chunk ;
		if ( a > m ) a = m ;
		if ( b > m ) b = m ;
		u32 nv = 0 ;
		for ( int q = a ; q < b ; ++q )
		{
			u64 code ;
			if ( t4_kc_canonical( sm->read, q, P.k, &code ) )
			{
				int c = (int)kc_lookup( P, code ) ;
				if ( c <= 0 )
					c = 1 ; // KmerCount.hpp:222-223
				sm->valid[q] = (u32)c ;
				++nv ;
			}
			else
				sm->valid[q] = 0 ;
		}
		u32 total ;
		u32 o = kc_scan( cx, nv, total ) ;
		for ( int q = a ; q < b ; ++q )
			if ( sm->valid[q] )
				sm->c[o++] = (int)sm->valid[q] ;
		T4_KC_SYNC() ;
		const int kk = (int)total ;
		int32_t *newLen = t4_x<int32_t>( P.newLen ) ;
		if ( kk == 0 )
		{
			if ( cx.tid == 0 )
			{
				minCnt[r] = medianCnt[r] = -len ; // KmerCount.hpp:229-239
				avgCnt[r] = (float)( -len ) ;
				if ( newLen )
					newLen[r] = P.qual ? 0 : len ; // `if ( qual != NULL ) read[0] = '\0'`
			}
			continue ;
		}
		// quality trimming (KmerCount.hpp:241-271), serial: the tail behind the last k-mer seen more than once is scanned from
		// the end; the read is cut at the leftmost position where the bad bases (Phred <= 15) reach 10 % of the tail
		int kUse = kk, trimStart = -1 ;
		if ( P.qual )
		{
			if ( cx.tid == 0 )
			{
				const char *q = t4_x<char>( P.qual ) + t4_x<u64>( P.seqOff )[r] ;
				int i ;
				for ( i = kk - 1 ; i >= 0 ; --i )
					if ( sm->c[i] > 1 )
						break ;
				++i ;
				int badCnt = 0, ts = -1 ;
				for ( int j = len - 1 ; j >= i + P.k - 1 ; --j )
					if ( q[j] - 32 <= 15 )
					{
						++badCnt ;
						if ( badCnt >= 0.1 * ( len - j ) )
							ts = j ;
					}
				int ku = kk ;
				if ( ts > 0 )
					ku = ts - P.k + 1 ;
				if ( ts > 0 && ts < P.k )
					ku = 0 ;
				if ( ku > kk )
					ku = kk ; // only with N's in a trimmed read: the reference then sorts stale entries of its shared buffer (undefined)
				sm->bi[2] = ku ;
				sm->bi[3] = ts ;
			}
			T4_KC_SYNC() ;
			kUse = sm->bi[2] ;
			trimStart = sm->bi[3] ;
			T4_KC_SYNC() ;
		}
		// min and the element of sorted rank kUse / 2 over the first kUse counts; the sum runs over ALL counts (the reference
		// sums before it trims, KmerCount.hpp:224, 274)
		int mn = 0x7fffffff ;
		u32 sum = 0 ;
		int med = -1 ;
		for ( int i = cx.tid ; i < kk ; i += cx.nt )
		{
			const int v = sm->c[i] ;
			sum += (u32)v ;
			if ( i >= kUse )
				continue ;
			if ( v < mn )
				mn = v ;
			int rank = 0 ;
			for ( int j = 0 ; j < kUse ; ++j )
			{
				const int w = sm->c[j] ;
				if ( w < v || ( w == v && j < i ) )
					++rank ;
			}
			if ( rank == kUse / 2 )
				med = v ;
		}
		// CTA reductions through shared memory (one slot per thread)
		T4_KC_SYNC() ;
		sm->scan[cx.tid] = (u32)mn ;
		T4_KC_SYNC() ;
		if ( cx.tid == 0 )
		{
			int x = 0x7fffffff ;
			for ( int t = 0 ; t < cx.nt ; ++t )
				if ( (int)sm->scan[t] < x )
					x = (int)sm->scan[t] ;
			sm->bi[0] = x ;
		}
		T4_KC_SYNC() ;
		u32 sumTotal ;
		kc_scan( cx, sum, sumTotal ) ;
		if ( med >= 0 )
			sm->bi[1] = med ; // exactly one thread holds the rank kUse / 2 element
		T4_KC_SYNC() ;
		if ( cx.tid == 0 )
		{
			const bool dropped = trimStart > 0 && trimStart < P.k ; // read[0] = '\0': the driver discards the read
			int minCount = kUse > 0 ? sm->bi[0] : sm->c[0] ;          // std::sort over nothing: c[0] is the first count as it stands
			const int medianCount = kUse > 0 ? sm->bi[1] : sm->c[0] ;
			for ( int i = 0 ; i < len ; ++i ) // KmerCount.hpp:275-283: over the ORIGINAL length; the cut only overwrote read[trimStart] (and read[0])
			{
				if ( ( trimStart > 0 && i == trimStart ) || ( dropped && i == 0 ) )
					continue ;
				if ( sm->read[i] == 'N' )
				{
					if ( minCount >= 0 )
						minCount = 0 ;
					else if ( minCount <= 0 )
						--minCount ;
				}
			}
			minCnt[r] = minCount ;
			medianCnt[r] = medianCount ;
			avgCnt[r] = (float)( (int)sumTotal / (double)kUse ) ; // `avgCount = sum / (double)k` into a float, KmerCount.hpp:274 (k = 0: inf)
			if ( newLen )
				newLen[r] = trimStart > 0 ? ( dropped ? 0 : trimStart ) : len ;
		}
	}
}

#endif
