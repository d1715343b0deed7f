// Host-side read sharding: which stream (= independent SeqSet) assembles which records of the driver's
// sorted read list.  Plain C++, no device code; used by t4_shard_reads (include/trust4_b200.h) and through it by the batch
// drop-in (integration/t4_seqset_adapter.hpp).  The Python twin for bench.py is trust4_b200/synth.py::shard_workload.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <queue>
#include <vector>

#include "../../include/trust4_b200.h"

namespace t4shard
{
// Predicted stream-engine cost of one record in microseconds on a B200, by the read's minimum 21-mer count (the abundance
// figure the driver sorts by): the row means of synth.COST_TABLE, which bench/fit_cost_model.py fits to measured stream
// cycles.  A duplicate record is one RepeatAddRead.
static const int kCostEdges[7] = { 1, 2, 4, 8, 16, 64, 256 } ;
static const double kCostRow[8] = { 352, 306, 306, 322, 356, 273, 209, 123 } ;
static const double kCostDup = 2.5 ;

static inline double RecordCost( const t4_read_desc &r )
{
	if ( r.flags & T4_RD_DUP )
		return kCostDup ;
	int b = 0 ;
	while ( b < 7 && r.min_cnt > kCostEdges[b] )
		++b ;
	return kCostRow[b] ;
}

// Contiguous blocks of the sorted list whose predicted costs are equal; a cut never falls inside a run of identical read
// strings (RepeatAddRead must see its first copy) nor, with barcodes, inside a barcode (barcodes are independent
// assemblies: main.cpp:1846-1859 purges them one by one).
static void RankStreams( const t4_read_desc *d, int64_t n, int S, bool wholeBarcodes, std::vector<int32_t> &streamOf )
{
	std::vector<double> cum( n ) ;
	double acc = 0 ;
	for ( int64_t i = 0 ; i < n ; ++i )
		cum[i] = ( acc += RecordCost( d[i] ) ) ;
	int64_t prev = 0 ;
	int s = 0 ;
	for ( int c = 1 ; c <= S ; ++c )
	{
		int64_t b = n ;
		if ( c < S )
		{
			b = std::lower_bound( cum.begin(), cum.end(), acc * c / S ) - cum.begin() ;
			if ( b < n )
				b = d[b].eq_lo ;
			if ( wholeBarcodes )
				while ( b > 0 && b < n && d[b].barcode == d[b - 1].barcode && d[b].barcode != -1 )
					--b ;
		}
		if ( b <= prev )
			continue ;
		for ( int64_t i = prev ; i < b ; ++i )
			streamOf[i] = s ;
		++s ;
		prev = b ;
	}
}

// Runs grouped by the gene of their rough annotation (names[name_id] of the run's first record: what InputNovelRead would
// call a contig seeded by it, mostly the V gene; -1 = reads that cannot seed), then: with t the smallest cap for which
// everything fits into S streams, a group dearer than t is cut into ceil(cost / t) contiguous sub-blocks, the others are
// packed largest-first into the least loaded of the remaining streams.  A clonotype's reads then meet in one SeqSet
// whatever their abundance rank; the reference shards --repseq data the same way (V-gene pseudo barcodes, main.cpp:1224-1235).
static void GeneStreams( const t4_read_desc *d, int64_t n, int S, std::vector<int32_t> &streamOf )
{
	int maxId = -1 ;
	for ( int64_t i = 0 ; i < n ; ++i )
		maxId = std::max( maxId, (int)d[d[i].eq_lo].name_id ) ;
	const int G = maxId + 2 ; // group g = name_id + 1
	std::vector<double> gcost( G, 0.0 ), cost( n ) ;
	std::vector<int32_t> grp( n ) ;
	double total = 0 ;
	for ( int64_t i = 0 ; i < n ; ++i )
	{
		grp[i] = d[d[i].eq_lo].name_id + 1 ;
		cost[i] = RecordCost( d[i] ) ;
		gcost[grp[i]] += cost[i] ;
		total += cost[i] ;
	}
	// streams needed under cap t
	auto need = [&]( double t, std::vector<int> &k, int &bins ) {
		k.assign( G, 0 ) ;
		double rest = 0 ;
		int64_t sum = 0 ;
		bool anySmall = false ;
		for ( int g = 0 ; g < G ; ++g )
		{
			if ( gcost[g] <= 0 )
				continue ;
			if ( gcost[g] > t )
				sum += ( k[g] = (int)std::ceil( gcost[g] / t ) ) ;
			else
			{
				rest += gcost[g] ;
				anySmall = true ;
			}
		}
		bins = anySmall ? std::max( 1, (int)std::ceil( 1.03 * rest / t ) ) : 0 ;
		return sum + bins ;
	} ;
	std::vector<int> k ;
	int bins = 0 ;
	double lo = total / S, hi = 2 * lo + *std::max_element( gcost.begin(), gcost.end() ) / S ;
	while ( need( hi, k, bins ) > S )
		hi *= 2 ;
	for ( int it = 0 ; it < 40 ; ++it )
	{
		double mid = 0.5 * ( lo + hi ) ;
		if ( need( mid, k, bins ) <= S )
			hi = mid ;
		else
			lo = mid ;
	}
	need( hi, k, bins ) ;
	// big groups: sub-blocks in sorted order, cut at run starts
	std::vector<int> first( G, -1 ) ;
	int nxt = 0 ;
	for ( int g = 0 ; g < G ; ++g )
		if ( k[g] > 0 )
		{
			first[g] = nxt ;
			nxt += k[g] ;
		}
	std::vector<double> seen( G, 0.0 ) ;
	std::vector<int> sub( G, 0 ) ;
	for ( int64_t i = 0 ; i < n ; ++i )
	{
		const int g = grp[i] ;
		if ( k[g] == 0 )
			continue ;
		if ( d[i].eq_lo == i ) // a run start may open the next sub-block
		{
			int want = (int)( seen[g] * k[g] / gcost[g] ) ;
			if ( want > k[g] - 1 )
				want = k[g] - 1 ;
			if ( want > sub[g] )
				sub[g] = want ;
		}
		seen[g] += cost[i] ;
		streamOf[i] = first[g] + sub[g] ;
	}
	// small groups: largest first into the least loaded bin
	std::vector<int> small ;
	for ( int g = 0 ; g < G ; ++g )
		if ( gcost[g] > 0 && k[g] == 0 )
			small.push_back( g ) ;
	std::sort( small.begin(), small.end(), [&]( int a, int b ) { return gcost[a] != gcost[b] ? gcost[a] > gcost[b] : a < b ; } ) ;
	typedef std::pair<double, int> Load ;
	std::priority_queue<Load, std::vector<Load>, std::greater<Load> > heap ;
	for ( int b = 0 ; b < bins ; ++b )
		heap.push( Load( 0.0, b ) ) ;
	std::vector<int> binOf( G, -1 ) ;
	for ( size_t j = 0 ; j < small.size() ; ++j )
	{
		Load l = heap.top() ;
		heap.pop() ;
		binOf[small[j]] = l.second ;
		heap.push( Load( l.first + gcost[small[j]], l.second ) ) ;
	}
	for ( int64_t i = 0 ; i < n ; ++i )
		if ( k[grp[i]] == 0 )
			streamOf[i] = nxt + binOf[grp[i]] ;
}

// See t4_shard_reads in include/trust4_b200.h.
static int Shard( t4_read_desc *d, int64_t n, int nStreams, int mode, int64_t *off, int64_t *order )
{
	if ( n < 0 || nStreams < 1 || !d || !off || !order )
		return -1 ;
	if ( nStreams > n )
		nStreams = n > 0 ? (int)n : 1 ;
	for ( int64_t i = 0 ; i < n ; ++i ) // the run fields are trusted below: check them once
		if ( d[i].eq_lo < 0 || d[i].eq_lo > i || d[i].eq_hi <= i || d[i].eq_hi > n || d[d[i].eq_lo].eq_lo != d[i].eq_lo
			|| d[i].mate_idx >= n )
			return -1 ;
	std::vector<int32_t> streamOf( n, 0 ) ;
	if ( n > 0 )
	{
		if ( mode == T4_SHARD_GENE )
			GeneStreams( d, n, nStreams, streamOf ) ;
		else
			RankStreams( d, n, nStreams, mode == T4_SHARD_BARCODE, streamOf ) ;
	}
	// dense stream ids (empty streams dropped), stable counting sort
	int maxS = 0 ;
	for ( int64_t i = 0 ; i < n ; ++i )
		maxS = std::max( maxS, (int)streamOf[i] ) ;
	std::vector<int64_t> cnt( maxS + 2, 0 ) ;
	for ( int64_t i = 0 ; i < n ; ++i )
		++cnt[streamOf[i]] ;
	std::vector<int> dense( maxS + 1, -1 ) ;
	int S = 0 ;
	for ( int s = 0 ; s <= maxS ; ++s )
		if ( cnt[s] > 0 )
			dense[s] = S++ ;
	if ( S == 0 )
		S = 1 ;
	std::vector<int64_t> start( S + 1, 0 ) ;
	for ( int s = 0 ; s <= maxS ; ++s )
		if ( dense[s] >= 0 )
			start[dense[s] + 1] = cnt[s] ;
	for ( int s = 0 ; s < S ; ++s )
		start[s + 1] += start[s] ;
	for ( int s = 0 ; s <= S ; ++s )
		off[s] = start[s] ;
	std::vector<int64_t> newPos( n ) ;
	{
		std::vector<int64_t> fill( start.begin(), start.end() - 1 ) ;
		for ( int64_t i = 0 ; i < n ; ++i )
		{
			const int s = dense[streamOf[i]] ;
			streamOf[i] = s ;
			newPos[i] = fill[s]++ ;
			order[newPos[i]] = i ;
		}
	}
	// records in stream order; run and mate indices relative to the stream (a mate in another stream gives no hint).
	// The members of a run that landed in one stream are consecutive there and keep their order.
	std::vector<t4_read_desc> out( n ) ;
	for ( int64_t j = 0 ; j < n ; ++j )
	{
		const int64_t i = order[j] ;
		const int s = streamOf[i] ;
		out[j] = d[i] ;
		const int64_t m = d[i].mate_idx ;
		out[j].mate_idx = ( m >= 0 && streamOf[m] == s ) ? (int32_t)( newPos[m] - off[s] ) : -1 ;
		const bool cont = j > off[s] && d[order[j - 1]].eq_lo == d[i].eq_lo ;
		out[j].eq_lo = cont ? out[j - 1].eq_lo : (int32_t)( j - off[s] ) ;
	}
	for ( int s = S - 1 ; s >= 0 ; --s )
		for ( int64_t j = off[s + 1] - 1 ; j >= off[s] ; --j )
		{
			const bool cont = j + 1 < off[s + 1] && out[j + 1].eq_lo == out[j].eq_lo ;
			out[j].eq_hi = cont ? out[j + 1].eq_hi : (int32_t)( j + 1 - off[s] ) ;
		}
	if ( n > 0 )
		memcpy( d, out.data(), sizeof( t4_read_desc ) * (size_t)n ) ;
	return S ;
}
} // namespace t4shard
