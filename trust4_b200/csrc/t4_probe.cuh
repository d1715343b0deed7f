// Grid-wide k-mer probe over FROZEN sets: SeqSet::GetHitsFromRead (SeqSet.hpp:1341-1501) + KmerIndex::Search
// (KmerIndex.hpp:104-116) for a whole batch of reads, one WARP per read, persistent CTAs.
//
// The stream kernel (t4_engine.h, one CTA per SeqSet) has to run its reads in order because AddRead mutates the set.
// A read-only pass has no such constraint: every read of the batch is independent, so the probe is laid out for
// memory-level parallelism instead of for serial latency --
//   * reads arrive 2-bit packed (t4_common.h) and are visited in (set, length-bucket) order (t4_bucket_kernel): the
//     warps of a CTA see equal work and consecutive warps hit the same directory, which keeps it in L2;
//   * a k-mer is one funnel shift of two packed words; each lane owns up to 9 positions of the read and issues
//     their directory probes back to back, one 256-bit load (LDG.E.256 = exactly the 32-byte T4Dir sector) each;
//   * hit slots are a warp prefix sum over the postings counts; the >= 100-postings skip rule and the stale
//     prevKmerCode quirk (SeqSet.hpp:1376-1392) only exist when some list has >= 100 postings -- then the warp runs
//     the reference's state machine over shuffled (code, size) pairs instead;
//   * postings: lists of <= 4 entries are one sector (T4_ALIGN 32) and are fetched by their owner lane with one
//     256-bit load; longer lists are pulled into shared memory by the TMA engine (cp.async.bulk + mbarrier, every
//     owner lane issues its own copies, all of a read's lists in flight at once) and converted to hit keys by the
//     whole warp with coalesced 8-byte stores; lists beyond the staging tile stream through 128-bit loads;
//   * the output range of a read is reserved with one atomicAdd, so hits of all reads form one dense key array.
// The keys are bit-identical to what c_get_hits() writes for the same read (same layout, same order).
#ifndef T4_PROBE_CUH
#define T4_PROBE_CUH

#if T4_CUDA

#define T4P_WARPS 8                 // warps per CTA
#define T4P_PMAX 9                  // positions per lane and tile
#define T4P_G 3                     // positions whose loads are issued together (groups of the unrolled loops)
#define T4P_TILE ( 32 * T4P_PMAX )  // positions (both strand passes) per tile: a 150 bp read at k = 9 has 284
#define T4P_STG 512                 // TMA staging tile per warp, postings (4 KB)
#define T4P_SHORT 4                 // a list of <= 4 postings is one 32-byte sector
#define T4P_TMA_MAX 256             // longer lists stream through 128-bit loads instead of the staging tile
#define T4P_NONE 0xffffffffu

struct T4ProbeParams
{
	char *A ;                  // arena
	const u64 *streamOff ;     // [nSets] arena offsets of the T4Stream records
	const t4_read_desc *descs ;
	const u64 *packed ;        // packed read pool: read i at packed + i * packStride
	u64 packStride ;
	const u64 *ord ;           // [nReads] read index | set index << 32, in (set, length bucket) order
	i64 nReads ;
	u64 *keys ;                // hit keys, dense
	u64 keyCap ;
	u64 *hitOff ;              // [nReads] first key of read i
	u32 *hitCnt ;              // [nReads] number of keys (invalid keys of the barcode filter included)
	u32 *hitFlags ;            // [nReads] bit 0: a k-mer has > 10000 postings (SeqSet.hpp:799), bit 1: serial rule path taken
	u64 *ctrl ;                // [0] read cursor [1] key cursor [2] overflow [3] lookups [4] postings [5] hits [6] packed read bytes [7] unsupported reads
	int allowTotalSkip ;
} ;

struct __align__( 16 ) T4ProbeWarp // sizeof is a multiple of 16: every warp's staging tile is a legal TMA destination
{
	u64 stg[T4P_STG] ;         // 16-byte aligned TMA destination
	u64 fw[18], rc[18] ;       // packed words (+ zero padding for the two-word funnel shift)
	u32 nm[20] ;
	u64 bar ;                  // mbarrier
} ;

__device__ __forceinline__ u64 t4p_extract( const u64 *W, int q, int k )
{
	const int w = q >> 5, s = ( q & 31 ) * 2 ;
	const u64 hi = W[w], lo = W[w + 1] ;
	const u64 v = s ? ( ( hi << s ) | ( lo >> ( 64 - s ) ) ) : hi ;
	return v >> ( 64 - 2 * k ) ;
}

__device__ __forceinline__ bool t4p_has_n( const u32 *M, int p, int k )
{
	const int w = p >> 5, s = p & 31 ;
	u32 v = M[w] >> s ;
	if ( s )
		v |= M[w + 1] << ( 32 - s ) ;
	return ( v & ( ( 1u << k ) - 1u ) ) != 0 ;
}

__device__ __forceinline__ void t4p_ld256( const void *p, u64 &a, u64 &b, u64 &c, u64 &d )
{
	asm volatile( "ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"( a ), "=l"( b ), "=l"( c ), "=l"( d ) : "l"( p ) ) ;
}
__device__ __forceinline__ void t4p_ld128( const void *p, u64 &a, u64 &b )
{
	asm volatile( "ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"( a ), "=l"( b ) : "l"( p ) ) ;
}

__device__ __forceinline__ u32 t4p_smem( const void *p ) { return (u32)__cvta_generic_to_shared( p ) ; }

__device__ __forceinline__ void t4p_bar_init( u64 *bar )
{
	asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"( t4p_smem( bar ) ) ) ;
	asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" ) ;
}
__device__ __forceinline__ void t4p_bar_expect( u64 *bar, u32 bytes )
{
	asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( t4p_smem( bar ) ), "r"( bytes ) : "memory" ) ;
}
__device__ __forceinline__ void t4p_bar_wait( u64 *bar, u32 parity )
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"T4P_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra T4P_DONE;\n"
		"bra T4P_WAIT;\n"
		"T4P_DONE:\n"
		"}\n" ::"r"( t4p_smem( bar ) ),
		"r"( parity )
		: "memory" ) ;
}
// TMA bulk copy global -> shared (non-tensor form): 16-byte aligned source / destination, size a multiple of 16
__device__ __forceinline__ void t4p_bulk_g2s( void *dst, const void *src, u32 bytes, u64 *bar )
{
	asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( t4p_smem( dst ) ),
		"l"( src ), "r"( bytes ), "r"( t4p_smem( bar ) )
		: "memory" ) ;
}

__device__ __forceinline__ u32 t4p_warp_excl_scan( u32 v, u32 &total, int lane )
{
	u32 inc = v ;
#pragma unroll
	for ( int d = 1 ; d < 32 ; d <<= 1 )
	{
		u32 t = __shfl_up_sync( 0xffffffffu, inc, d ) ;
		if ( lane >= d )
			inc += t ;
	}
	total = __shfl_sync( 0xffffffffu, inc, 31 ) ;
	return inc - v ;
}

// per-read constants
struct T4ProbeRead
{
	const T4Dir *dir ;
	u32 dirMask ;
	const T4Contig *seqs ;
	u64 salt ;                 // barcode salt of the directory key (t4_index_key)
	int k, len, m, strand, barcode ;
} ;

// Serial rule state (SeqSet.hpp:1376-1392, 1441-1455), carried across tiles and from the forward into the reverse pass.
struct T4ProbeScan
{
	u64 prev ;
	int skipCnt, curPass ;
	u32 total, lookups ;
	int big ;
} ;

// Directory probes of one tile: position x = tile0 + c * 32 + lane -> cnt[c], lo[c] (0 postings when the k-mer holds an N,
// the pass is not probed for this strand, or the code is absent).  Returns true iff some list has >= 100 postings.
__device__ __forceinline__ bool t4p_probe_tile( const T4ProbeRead &R, const T4ProbeWarp *sw, const char *A, int tile0, int lane,
	u32 ( &cnt )[T4P_PMAX], u32 ( &lo )[T4P_PMAX] )
{
	bool large = false ;
	// groups of T4P_G positions: the first-slot loads of a group are in flight together
#pragma unroll
	for ( int h = 0 ; h < T4P_PMAX / T4P_G ; ++h )
	{
		u64 key[T4P_G], v[T4P_G][4] ;
		u32 slot[T4P_G] ;
#pragma unroll
		for ( int cc = 0 ; cc < T4P_G ; ++cc )
		{
			const int c = h * T4P_G + cc ;
			const int x = tile0 + c * 32 + lane ;
			cnt[c] = 0 ;
			lo[c] = 0 ;
			key[cc] = 0 ;
			if ( x >= 2 * R.m )
				continue ;
			const int pass = x >= R.m ;
			const int q = pass ? x - R.m : x ;
			if ( ( pass == 0 && R.strand == -1 ) || ( pass == 1 && R.strand == 1 ) )
				continue ;
			// validity window: forward positions [q, q + k) or, for the reverse pass, [len - q - k, len - q)
			if ( t4p_has_n( sw->nm, pass ? R.len - q - R.k : q, R.k ) )
				continue ;
			const u64 code = t4p_extract( pass ? sw->rc : sw->fw, q, R.k ) ;
			key[cc] = code + R.salt + 1 ;
			slot[cc] = (u32)( ( key[cc] * 0x9E3779B97F4A7C15ull ) >> 32 ) & R.dirMask ;
			t4p_ld256( R.dir + slot[cc], v[cc][0], v[cc][1], v[cc][2], v[cc][3] ) ;
		}
#pragma unroll
		for ( int cc = 0 ; cc < T4P_G ; ++cc )
		{
			const int c = h * T4P_G + cc ;
			if ( key[cc] == 0 )
				continue ;
			u64 kk = v[cc][0], lOff = v[cc][1], cw = v[cc][2], pad ;
			u32 s = slot[cc] ;
			while ( kk != key[cc] && kk != 0 ) // linear probing past a colliding slot (rare at load factor <= 1/2)
			{
				s = ( s + 1 ) & R.dirMask ;
				t4p_ld256( R.dir + s, kk, lOff, cw, pad ) ;
			}
			if ( kk == key[cc] )
			{
				cnt[c] = (u32)cw ;
				lo[c] = (u32)( lOff >> 5 ) ; // lists are 32-byte aligned (T4_ALIGN)
				if ( cnt[c] >= 100 )
					large = true ;
			}
		}
	}
	return __any_sync( 0xffffffffu, large ) ;
}

// No list reaches 100 postings: "taken" is a per-position predicate (first k-mer of the pass, or code differs from the
// previous k-mer's -- N counted as A, KmerCode.hpp:94), hit slots are an exclusive prefix sum in position order.
__device__ __forceinline__ void t4p_scan_fast( const T4ProbeRead &R, const T4ProbeWarp *sw, int tile0, int lane, const u32 ( &cnt )[T4P_PMAX],
	u32 ( &base )[T4P_PMAX], T4ProbeScan &S )
{
#pragma unroll
	for ( int c = 0 ; c < T4P_PMAX ; ++c )
	{
		const int x = tile0 + c * 32 + lane ;
		bool taken = false ;
		if ( x < 2 * R.m )
		{
			const int pass = x >= R.m ;
			const int q = pass ? x - R.m : x ;
			if ( !( ( pass == 0 && R.strand == -1 ) || ( pass == 1 && R.strand == 1 ) ) )
			{
				const u64 *W = pass ? sw->rc : sw->fw ;
				taken = ( q == 0 ) || ( t4p_extract( W, q, R.k ) != t4p_extract( W, q - 1, R.k ) ) ;
			}
		}
		const u32 v = taken ? cnt[c] : 0 ;
		u32 tot ;
		const u32 o = t4p_warp_excl_scan( v, tot, lane ) ;
		base[c] = ( taken && v > 0 ) ? S.total + o : T4P_NONE ;
		S.total += tot ;
		S.lookups += __popc( __ballot_sync( 0xffffffffu, taken ) ) ;
	}
}

// Some list has >= 100 postings: the reference's loop, position by position, warp-uniform over shuffled sizes.
__device__ __forceinline__ void t4p_scan_serial( const T4ProbeRead &R, const T4ProbeWarp *sw, int tile0, int lane, int allowTotalSkip,
	const u32 ( &cnt )[T4P_PMAX], u32 ( &base )[T4P_PMAX], T4ProbeScan &S )
{
	const int skipLimit = R.k / 2 ;
#pragma unroll
	for ( int c = 0 ; c < T4P_PMAX ; ++c )
	{
		base[c] = T4P_NONE ;
		const int x0 = tile0 + c * 32 ;
		if ( x0 >= 2 * R.m )
			continue ;
		for ( int src = 0 ; src < 32 ; ++src )
		{
			const int x = x0 + src ;
			const u32 size = __shfl_sync( 0xffffffffu, cnt[c], src ) ;
			if ( x >= 2 * R.m )
				break ;
			const int pass = x >= R.m ;
			const int q = pass ? x - R.m : x ;
			if ( ( pass == 0 && R.strand == -1 ) || ( pass == 1 && R.strand == 1 ) )
				continue ;
			if ( pass != S.curPass )
			{
				S.curPass = pass ;
				S.skipCnt = 0 ;
			}
			const u64 code = t4p_extract( pass ? sw->rc : sw->fw, q, R.k ) ;
			const int i = q + R.k - 1 ;
			if ( i == R.k - 1 || code != S.prev )
			{
				++S.lookups ;
				if ( size >= 100 && i != R.k - 1 && i != R.len - 1 && S.skipCnt < skipLimit )
				{
					++S.skipCnt ;
					continue ; // prevKmerCode keeps its stale value (SeqSet.hpp:1381-1388)
				}
				if ( size >= 100 && allowTotalSkip )
					continue ;
				S.skipCnt = 0 ;
				if ( size > 0 )
				{
					if ( lane == src )
						base[c] = S.total ;
					S.total += size ;
					if ( R.barcode == -1 && size > T4_BIG_REPEAT )
						S.big = 1 ;
				}
			}
			S.prev = code ;
		}
	}
}

__device__ __forceinline__ u64 t4p_key( const T4ProbeRead &R, int pass, int q, u64 posting, int big )
{
	const int idx = (int)( posting >> 32 ) ;
	const int off = (int)(u32)posting ;
	u64 key = t4_key_of( pass ? -1 : 1, idx, q, off, big ) ;
	if ( R.barcode != -1 && __ldg( &R.seqs[idx].barcode ) != R.barcode ) // SeqSet.hpp:1418
		key = T4_KEY_INVALID ;
	return key ;
}

// Postings -> hit keys for one tile.  out: first key of this read.
__device__ __forceinline__ void t4p_emit_tile( const T4ProbeRead &R, T4ProbeWarp *sw, const char *A, int tile0, int lane, const u32 ( &cnt )[T4P_PMAX],
	const u32 ( &lo )[T4P_PMAX], const u32 ( &base )[T4P_PMAX], u64 *out, u32 &barPhase )
{
	// ---- lists of <= 4 postings: one sector, fetched by the owner lane (the loads of a group are issued before the first use)
#pragma unroll
	for ( int h = 0 ; h < T4P_PMAX / T4P_G ; ++h )
	{
		u64 p[T4P_G][4] ;
#pragma unroll
		for ( int cc = 0 ; cc < T4P_G ; ++cc )
		{
			const int c = h * T4P_G + cc ;
			if ( base[c] != T4P_NONE && cnt[c] <= T4P_SHORT )
				t4p_ld256( A + ( (u64)lo[c] << 5 ), p[cc][0], p[cc][1], p[cc][2], p[cc][3] ) ;
		}
#pragma unroll
		for ( int cc = 0 ; cc < T4P_G ; ++cc )
		{
			const int c = h * T4P_G + cc ;
			if ( base[c] != T4P_NONE && cnt[c] <= T4P_SHORT )
			{
				const int x = tile0 + c * 32 + lane ;
				const int pass = x >= R.m ;
				const int q = pass ? x - R.m : x ;
				u64 *o = out + base[c] ;
#pragma unroll
				for ( int j = 0 ; j < T4P_SHORT ; ++j )
					if ( j < (int)cnt[c] )
						o[j] = t4p_key( R, pass, q, p[cc][j], 0 ) ;
			}
		}
	}
	// ---- lists of 5 .. T4P_TMA_MAX postings: TMA into the staging tile, rounds of at most T4P_STG postings
	u32 pending = 0 ; // bit c: list of position c still to be staged
#pragma unroll
	for ( int c = 0 ; c < T4P_PMAX ; ++c )
		if ( base[c] != T4P_NONE && cnt[c] > T4P_SHORT && cnt[c] <= T4P_TMA_MAX )
			pending |= 1u << c ;
	while ( __any_sync( 0xffffffffu, pending != 0 ) )
	{
		// staging offsets of this round: exclusive prefix (in position order) over the even-rounded counts of the pending lists;
		// a list is taken iff it fits entirely, which selects a prefix of the pending lists
		u32 sb[T4P_PMAX] ;
		u32 run = 0 ;
		bool full = false ;
#pragma unroll
		for ( int c = 0 ; c < T4P_PMAX ; ++c )
		{
			const bool pend = ( pending >> c ) & 1u ;
			const u32 v = pend ? ( ( cnt[c] + 1 ) & ~1u ) : 0 ;
			u32 tot ;
			const u32 o = t4p_warp_excl_scan( v, tot, lane ) ;
			sb[c] = T4P_NONE ;
			if ( pend && !full && run + o + v <= T4P_STG )
				sb[c] = run + o ;
			// once one pending list does not fit, no later list may be taken (keeps the selection a prefix)
			const u32 miss = __ballot_sync( 0xffffffffu, pend && sb[c] == T4P_NONE ) ;
			if ( miss )
			{
				const int first = __ffs( miss ) - 1 ;
				if ( lane > first && sb[c] != T4P_NONE )
					sb[c] = T4P_NONE ;
				full = true ;
			}
			run += tot ;
		}
		u32 myBytes = 0 ;
#pragma unroll
		for ( int c = 0 ; c < T4P_PMAX ; ++c )
			if ( sb[c] != T4P_NONE )
				myBytes += ( ( cnt[c] + 1 ) & ~1u ) * 8 ;
		u32 roundBytes = myBytes ;
#pragma unroll
		for ( int d = 16 ; d > 0 ; d >>= 1 )
			roundBytes += __shfl_xor_sync( 0xffffffffu, roundBytes, d ) ;
		// the previous round's generic-proxy reads of the tile are ordered before the async-proxy writes of this one
		asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ) ;
		__syncwarp() ;
		if ( lane == 0 )
			t4p_bar_expect( &sw->bar, roundBytes ) ;
		__syncwarp() ;
#pragma unroll
		for ( int c = 0 ; c < T4P_PMAX ; ++c )
			if ( sb[c] != T4P_NONE )
				t4p_bulk_g2s( sw->stg + sb[c], A + ( (u64)lo[c] << 5 ), ( ( cnt[c] + 1 ) & ~1u ) * 8, &sw->bar ) ;
		t4p_bar_wait( &sw->bar, barPhase ) ;
		barPhase ^= 1 ;
		// convert: one staged list at a time, the whole warp on it
#pragma unroll
		for ( int c = 0 ; c < T4P_PMAX ; ++c )
		{
			u32 mask = __ballot_sync( 0xffffffffu, sb[c] != T4P_NONE ) ;
			while ( mask )
			{
				const int src = __ffs( mask ) - 1 ;
				mask &= mask - 1 ;
				const u32 n = __shfl_sync( 0xffffffffu, cnt[c], src ) ;
				const u32 so = __shfl_sync( 0xffffffffu, sb[c], src ) ;
				const u32 bo = __shfl_sync( 0xffffffffu, base[c], src ) ;
				const int x = tile0 + c * 32 + src ;
				const int pass = x >= R.m ;
				const int q = pass ? x - R.m : x ;
				for ( u32 j = lane ; j < n ; j += 32 )
					out[bo + j] = t4p_key( R, pass, q, sw->stg[so + j], 0 ) ;
			}
			if ( sb[c] != T4P_NONE )
				pending &= ~( 1u << c ) ;
		}
		__syncwarp() ;
	}
	// ---- longer lists: streamed by the whole warp, two postings (128 bits) per lane and load
#pragma unroll
	for ( int c = 0 ; c < T4P_PMAX ; ++c )
	{
		u32 mask = __ballot_sync( 0xffffffffu, base[c] != T4P_NONE && cnt[c] > T4P_TMA_MAX ) ;
		while ( mask )
		{
			const int src = __ffs( mask ) - 1 ;
			mask &= mask - 1 ;
			const u32 n = __shfl_sync( 0xffffffffu, cnt[c], src ) ;
			const u64 l = (u64)__shfl_sync( 0xffffffffu, lo[c], src ) << 5 ;
			const u32 bo = __shfl_sync( 0xffffffffu, base[c], src ) ;
			const int x = tile0 + c * 32 + src ;
			const int pass = x >= R.m ;
			const int q = pass ? x - R.m : x ;
			const int big = ( R.barcode == -1 && n > T4_BIG_REPEAT ) ? 1 : 0 ;
			for ( u32 j = 2 * lane ; j < n ; j += 64 )
			{
				u64 a, b ;
				t4p_ld128( A + l + 8ull * j, a, b ) ;
				out[bo + j] = t4p_key( R, pass, q, a, big ) ;
				if ( j + 1 < n )
					out[bo + j + 1] = t4p_key( R, pass, q, b, big ) ;
			}
		}
	}
}

__global__ void __launch_bounds__( 32 * T4P_WARPS, 2 ) t4_probe_kernel( T4ProbeParams P )
{
	__shared__ __align__( 16 ) T4ProbeWarp smem[T4P_WARPS] ;
	const int lane = threadIdx.x & 31 ;
	T4ProbeWarp *sw = &smem[threadIdx.x >> 5] ;
	if ( lane == 0 )
		t4p_bar_init( &sw->bar ) ;
	if ( lane < 2 )
	{
		sw->fw[16 + lane] = 0 ;
		sw->rc[16 + lane] = 0 ;
	}
	__syncwarp() ;
	u32 barPhase = 0 ;
	u64 accLook = 0, accPost = 0, accBytes = 0, accUnsup = 0 ;
	while ( 1 )
	{
		unsigned long long w = 0 ;
		if ( lane == 0 )
			w = atomicAdd( (unsigned long long *)&P.ctrl[0], 1ull ) ;
		w = __shfl_sync( 0xffffffffu, w, 0 ) ;
		if ( (i64)w >= P.nReads )
			break ;
		const u64 ent = __ldg( P.ord + w ) ;
		const u32 ri = (u32)ent, si = (u32)( ent >> 32 ) ;
		const t4_read_desc *d = P.descs + ri ;
		const T4Stream *st = (const T4Stream *)( P.A + __ldg( P.streamOff + si ) ) ;
		T4ProbeRead R ;
		R.len = d->len ;
		R.barcode = d->barcode ;
		R.strand = d->strand_in ;
		R.k = st->kmerLength ;
		R.m = R.len - R.k + 1 ;
		R.dir = (const T4Dir *)( P.A + st->dirOff ) ;
		R.dirMask = st->dirCap - 1 ;
		R.seqs = (const T4Contig *)( P.A + st->seqsOff ) ;
		R.salt = st->considerBarcode ? ( (u64)(u32)( R.barcode + 1 ) << ( 2 * R.k ) ) : 0ull ;
		if ( R.len > T4_DEV_MAX_READ || R.len < R.k )
		{
			if ( lane == 0 )
			{
				P.hitOff[ri] = 0 ;
				P.hitCnt[ri] = 0 ;
				P.hitFlags[ri] = 0 ;
				if ( R.len > T4_DEV_MAX_READ )
					++accUnsup ;
			}
			continue ;
		}
		// packed read -> shared memory (<= 16 + 16 words + 16 mask words)
		{
			const int W = (int)t4_pack_w( R.len ) ;
			const u64 *pk = P.packed + (u64)ri * P.packStride ;
			__syncwarp() ;
			if ( lane < W )
			{
				sw->fw[lane] = __ldg( pk + lane ) ;
				sw->rc[lane] = __ldg( pk + W + lane ) ;
				sw->nm[lane] = __ldg( (const u32 *)( pk + 2 * W ) + lane ) ;
			}
			else if ( lane < 18 )
			{
				sw->fw[lane] = 0 ;
				sw->rc[lane] = 0 ;
				sw->nm[lane] = 0 ;
			}
			if ( lane >= 18 && lane < 20 )
				sw->nm[lane] = 0 ;
			__syncwarp() ;
		}
		T4ProbeScan S ;
		u32 cnt[T4P_PMAX], base[T4P_PMAX], lo[T4P_PMAX] ;
		const int nTiles = ( 2 * R.m + T4P_TILE - 1 ) / T4P_TILE ;
		u32 flags = 0 ;
		u64 *out = 0 ;
		// sweep 0 counts (directory probes + slot assignment), then the output range is reserved, sweep 1 emits.  A read
		// that fits one tile (<= 288 positions: 150 bp at k >= 7) keeps its probe results in registers between the two;
		// a longer one probes its tiles again (L1/L2 hits) and always uses the serial rules, which reduce to the plain
		// predicate when no list is large.
		for ( int sweep = 0 ; sweep < 2 ; ++sweep )
		{
			if ( sweep == 0 || nTiles > 1 )
			{
				S.prev = 0 ; S.skipCnt = 0 ; S.curPass = -1 ; S.total = 0 ; S.lookups = 0 ; S.big = 0 ;
			}
			for ( int t = 0 ; t < nTiles ; ++t )
			{
				if ( sweep == 0 || nTiles > 1 )
				{
					const bool large = t4p_probe_tile( R, sw, P.A, t * T4P_TILE, lane, cnt, lo ) ;
					if ( nTiles == 1 && !large )
						t4p_scan_fast( R, sw, 0, lane, cnt, base, S ) ;
					else
					{
						t4p_scan_serial( R, sw, t * T4P_TILE, lane, P.allowTotalSkip, cnt, base, S ) ;
						flags |= 2 ;
					}
				}
				if ( sweep == 1 )
					t4p_emit_tile( R, sw, P.A, t * T4P_TILE, lane, cnt, lo, base, out, barPhase ) ;
			}
			if ( sweep == 0 )
			{
				const u32 T = S.total ;
				unsigned long long o0 = 0 ;
				if ( lane == 0 )
					o0 = atomicAdd( (unsigned long long *)&P.ctrl[1], (unsigned long long)T ) ;
				o0 = __shfl_sync( 0xffffffffu, o0, 0 ) ;
				const bool fits = o0 + T <= P.keyCap ;
				if ( lane == 0 )
				{
					P.hitOff[ri] = o0 ;
					P.hitCnt[ri] = fits ? T : 0 ;
					P.hitFlags[ri] = flags | ( S.big ? 1u : 0u ) ;
					if ( !fits )
						P.ctrl[2] = 1 ;
				}
				accLook += S.lookups ;
				accPost += T ;
				accBytes += ( R.len + 3 ) / 4 ;
				if ( !fits || T == 0 )
					break ;
				out = P.keys + o0 ;
			}
		}
	}
	if ( lane == 0 )
	{
		if ( accLook ) atomicAdd( (unsigned long long *)&P.ctrl[3], (unsigned long long)accLook ) ;
		if ( accPost )
		{
			atomicAdd( (unsigned long long *)&P.ctrl[4], (unsigned long long)accPost ) ;
			atomicAdd( (unsigned long long *)&P.ctrl[5], (unsigned long long)accPost ) ;
		}
		if ( accBytes ) atomicAdd( (unsigned long long *)&P.ctrl[6], (unsigned long long)accBytes ) ;
		if ( accUnsup ) atomicAdd( (unsigned long long *)&P.ctrl[7], (unsigned long long)accUnsup ) ;
	}
}

// ASCII pool -> 2-bit packed pool (t4_common.h layout).  One thread per (read, word): 32 forward bases, 32
// reverse-complement bases and 32 mask bits.  *odd is set when a read holds a character outside ACGTN (the packed
// form cannot represent it; callers then keep using the ASCII pool for the assembly).
__global__ void t4_pack_reads_kernel( const t4_read_desc *descs, i64 n, const char *pool, u64 packStride, u64 *packed, u32 *odd )
{
	const i64 r = blockIdx.x ;
	if ( r >= n )
		return ;
	const int len = descs[r].len ;
	if ( len <= 0 || len > T4_DEV_MAX_READ )
		return ;
	const char *s = pool + descs[r].seq_off ;
	const int W = (int)t4_pack_w( len ) ;
	u64 *fw = packed + (u64)r * packStride, *rc = fw + W ;
	u32 *nm = (u32 *)( fw + 2 * W ) ;
	for ( int w = threadIdx.x ; w < W ; w += blockDim.x )
		t4_pack_word( s, len, w, fw + w, rc + w, nm + w, odd ) ;
	if ( threadIdx.x == 0 && ( W & 1 ) )
		nm[W] = 0 ;
}

// Visit order of the probe: reads of one set together, inside a set by length bucket (32 bases per bucket, longest
// first) so that the warps of a CTA carry equal numbers of k-mers.  One CTA per set.
__global__ void t4_bucket_kernel( const t4_read_desc *descs, const i64 *descOff, u64 *ord )
{
	__shared__ u32 hist[17], cur[17] ;
	const int s = blockIdx.x ;
	const i64 lo = descOff[s], hi = descOff[s + 1] ;
	if ( threadIdx.x < 17 )
		hist[threadIdx.x] = 0 ;
	__syncthreads() ;
	for ( i64 i = lo + threadIdx.x ; i < hi ; i += blockDim.x )
	{
		int b = 16 - min( 16, max( 0, descs[i].len ) >> 5 ) ;
		atomicAdd( &hist[b], 1u ) ;
	}
	__syncthreads() ;
	if ( threadIdx.x == 0 )
	{
		u32 a = 0 ;
		for ( int b = 0 ; b < 17 ; ++b )
		{
			cur[b] = a ;
			a += hist[b] ;
		}
	}
	__syncthreads() ;
	for ( i64 i = lo + threadIdx.x ; i < hi ; i += blockDim.x )
	{
		int b = 16 - min( 16, max( 0, descs[i].len ) >> 5 ) ;
		u32 p = atomicAdd( &cur[b], 1u ) ;
		ord[lo + p] = (u64)(u32)i | ( (u64)(u32)s << 32 ) ;
	}
}

#endif // T4_CUDA
#endif
