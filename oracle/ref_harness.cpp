// TEST INFRASTRUCTURE (oracle) -- not part of the product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the library built from this file.
//
// Library-level oracle: a thin C ABI over the UNMODIFIED
// reference classes, compiled from the sources where they lie in /root/reference
// (see oracle/Makefile; output oracle/_ref/libt4ref.so).  `#define private public`
// gives the stage dumps (hits, chains, overlaps, index) access to SeqSet's
// private helpers.  Nothing of the reference is copied into this repository.
//
// Besides the per-call wrappers, t4ref_run_descs() restates the stage-1 driver
// loop rules (main.cpp:1583-1880, rescue pass 1897-1940) over t4_read_desc
// records -- the oracle for the batch entry t4_seqset_add_reads_batch().  That
// restatement is pinned against the stock binary by tests/test_oracle.py
// (call-trace replay must reproduce trust4's own _raw.out byte for byte).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <time.h>
#include <assert.h>
#include <limits.h>
#include <map>
#include <algorithm>
#include <vector>
#include <string>

#define private public
#include "SeqSet.hpp"
#undef private

#include "KmerCount.hpp"

#include "../include/trust4_b200.h"

// Each reference executable defines these (main.cpp:39-44); values restated from defs.h's contract.
int nucToNum[26] = { 0, -1, 1, -1, -1, -1, 2,
	-1, -1, -1, -1, -1, -1, 0,
	-1, -1, -1, -1, -1, 3,
	-1, -1, -1, -1, -1, -1 } ;
char numToNuc[26] = {'A', 'C', 'G', 'T'} ;

extern "C" {

void *t4ref_create( int k ) { return new SeqSet( k ) ; }
void t4ref_destroy( void *h ) { delete (SeqSet *)h ; }
int t4ref_set_hit_len_required( void *h, int l ) { return ((SeqSet *)h)->SetHitLenRequired( l ) ; }
double t4ref_set_novel_seq_similarity( void *h, double v ) { return ((SeqSet *)h)->SetNovelSeqSimilarity( v ) ; }
void t4ref_set_consider_barcode_in_hash( void *h, int on ) { ((SeqSet *)h)->SetConsiderBarcodeInIndexHash( on != 0 ) ; }
void t4ref_set_is_long( void *h, int on ) { ((SeqSet *)h)->SetIsLongSeqSet( on != 0 ) ; }
int t4ref_size( void *h ) { return ((SeqSet *)h)->Size() ; }
int t4ref_kmer_length( void *h ) { return ((SeqSet *)h)->kmerLength ; }

int t4ref_add_read( void *h, const char *read, const char *geneName, int *strand, int barcode, int minKmerCount,
	int repetitive, double thr )
{
	char *r = strdup( read ) ;
	char name[64] ;
	strncpy( name, geneName, 63 ) ; name[63] = '\0' ;
	int s = *strand ;
	int ret = ((SeqSet *)h)->AddRead( r, name, s, barcode, minKmerCount, repetitive != 0, thr ) ;
	*strand = s ;
	free( r ) ;
	return ret ;
}

int t4ref_repeat_add_read( void *h, const char *read )
{
	char *r = strdup( read ) ;
	int ret = ((SeqSet *)h)->RepeatAddRead( r ) ;
	free( r ) ;
	return ret ;
}

int t4ref_input_novel_read( void *h, const char *id, const char *read, int strand, int barcode )
{
	char *r = strdup( read ) ;
	int ret = ((SeqSet *)h)->InputNovelRead( id, r, strand, barcode ) ;
	free( r ) ;
	return ret ;
}

void t4ref_update_all_consensus( void *h ) { ((SeqSet *)h)->UpdateAllConsensus() ; }
// SeqSet::ReleaseFinishedBarcodeSeq as the driver calls it (main.cpp:1855), ReleaseShallowContigs (main.cpp:1954),
// InputNovelFa (main.cpp:711)
void t4ref_release_finished_barcode( void *h, int barcode, int contigMinCov )
{
	std::map<int, int> fin ;
	fin[barcode] = 1 ;
	((SeqSet *)h)->ReleaseFinishedBarcodeSeq( fin, true, contigMinCov, true ) ;
}
void t4ref_release_shallow_contigs( void *h, int minCov ) { ((SeqSet *)h)->ReleaseShallowContigs( minCov ) ; }
void t4ref_input_novel_fa( void *h, const char *filename ) { ((SeqSet *)h)->InputNovelFa( (char *)filename ) ; }
int t4ref_num_read( void *h, int slot )
{
	SeqSet *s = (SeqSet *)h ;
	if ( slot < 0 || slot >= (int)s->seqs.size() || s->seqs[slot].consensus == NULL )
		return -1 ;
	return s->seqs[slot].numRead ;
}
void t4ref_change_kmer_length( void *h, int kl ) { ((SeqSet *)h)->ChangeKmerLength( kl ) ; }

int t4ref_has_motif( void *h, const char *read, int strand )
{
	char *r = strdup( read ) ;
	int ret = ((SeqSet *)h)->HasMotif( r, strand ) ;
	free( r ) ;
	return ret ;
}

void t4ref_reverse_complement_in_place( void *h, char *seq, int len ) { ((SeqSet *)h)->ReverseComplementInPlace( seq, len ) ; }

// SeqSet::Output into a malloc'ed buffer; caller frees with t4ref_free.
int t4ref_output_mem( void *h, char **buf, size_t *len )
{
	FILE *fp = open_memstream( buf, len ) ;
	if ( fp == NULL )
		return -1 ;
	((SeqSet *)h)->Output( fp, NULL ) ;
	fclose( fp ) ;
	return 0 ;
}
void t4ref_free( void *p ) { free( p ) ; }

// Contig accessors.
int t4ref_get_contig( void *h, int slot, char *consensus, int consensusCap, int32_t *posWeight, char *name, int nameCap,
	int *barcode, int *numRead, int *minLeft, int *minRight )
{
	SeqSet *s = (SeqSet *)h ;
	if ( slot < 0 || slot >= (int)s->seqs.size() || s->seqs[slot].consensus == NULL )
		return -1 ;
	struct _seqWrapper &seq = s->seqs[slot] ;
	int len = seq.consensusLen ;
	if ( consensus != NULL && consensusCap > len )
	{
		memcpy( consensus, seq.consensus, len ) ;
		consensus[len] = '\0' ;
	}
	if ( posWeight != NULL )
		for ( int i = 0 ; i < len ; ++i )
			for ( int j = 0 ; j < 4 ; ++j )
				posWeight[4 * i + j] = seq.posWeight[i].count[j] ;
	if ( name != NULL && nameCap > 0 )
	{
		strncpy( name, seq.name, nameCap - 1 ) ;
		name[nameCap - 1] = '\0' ;
	}
	if ( barcode ) *barcode = seq.barcode ;
	if ( numRead ) *numRead = seq.numRead ;
	if ( minLeft ) *minLeft = seq.minLeftExtAnchor ;
	if ( minRight ) *minRight = seq.minRightExtAnchor ;
	return len ;
}

// ---- stage dumps ----------------------------------------------------------
// GetHitsFromRead + SortHits (SeqSet.hpp:1341, 1306): int32[5] per hit.
int t4ref_get_hits( void *h, const char *read, int strand, int barcode, int allowTotalSkip, int32_t *out, int cap )
{
	SeqSet *s = (SeqSet *)h ;
	int len = strlen( read ) ;
	if ( len < s->kmerLength )
		return 0 ;
	char *r = strdup( read ) ;
	char *rc = new char[len + 1] ;
	SimpleVector<struct _hit> hits ;
	s->GetHitsFromRead( r, rc, strand, barcode, allowTotalSkip != 0, hits, NULL ) ;
	s->SortHits( hits, true ) ;
	int n = hits.Size() ;
	for ( int i = 0 ; i < n && i < cap ; ++i )
	{
		out[5 * i] = hits[i].indexHit.idx ;
		out[5 * i + 1] = hits[i].indexHit.offset ;
		out[5 * i + 2] = hits[i].readOffset ;
		out[5 * i + 3] = hits[i].strand ;
		out[5 * i + 4] = hits[i].repeats ;
	}
	delete[] rc ;
	free( r ) ;
	return n ;
}

// GetHitsFromRead + SortHits + GetOverlapsFromHits (SeqSet.hpp:763) as AddRead's
// full pass runs it (filter 1, conservativeChain false): int32[8] per chain =
// {seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt, nHitCoords}; the
// hitCoords (a,b) pairs are appended to coords (2 ints each) in chain order.
int t4ref_get_chains( void *h, const char *read, int strand, int barcode, int allowTotalSkip, int filter,
	int32_t *out, int cap, int32_t *coords, int coordCap, int *nCoords )
{
	SeqSet *s = (SeqSet *)h ;
	int len = strlen( read ) ;
	*nCoords = 0 ;
	if ( len < s->kmerLength )
		return 0 ;
	char *r = strdup( read ) ;
	char *rc = new char[len + 1] ;
	SimpleVector<struct _hit> hits ;
	std::vector<struct _overlap> overlaps ;
	s->GetHitsFromRead( r, rc, strand, barcode, allowTotalSkip != 0, hits, NULL ) ;
	s->SortHits( hits, true ) ;
	int n = s->GetOverlapsFromHits( hits, s->hitLenRequired, filter, false, overlaps ) ;
	int c = 0 ;
	for ( int i = 0 ; i < n ; ++i )
	{
		int m = overlaps[i].hitCoords->Size() ;
		if ( i < cap )
		{
			out[8 * i] = overlaps[i].seqIdx ;
			out[8 * i + 1] = overlaps[i].readStart ;
			out[8 * i + 2] = overlaps[i].readEnd ;
			out[8 * i + 3] = overlaps[i].seqStart ;
			out[8 * i + 4] = overlaps[i].seqEnd ;
			out[8 * i + 5] = overlaps[i].strand ;
			out[8 * i + 6] = overlaps[i].matchCnt ;
			out[8 * i + 7] = m ;
		}
		for ( int j = 0 ; j < m ; ++j )
		{
			if ( c + j < coordCap )
			{
				coords[2 * ( c + j )] = (*overlaps[i].hitCoords)[j].a ;
				coords[2 * ( c + j ) + 1] = (*overlaps[i].hitCoords)[j].b ;
			}
		}
		c += m ;
		delete overlaps[i].hitCoords ;
	}
	*nCoords = c ;
	delete[] rc ;
	free( r ) ;
	return n ;
}

// GetOverlapsFromRead (SeqSet.hpp:1508) with readType 0: int32[8] per overlap + similarity.
int t4ref_get_overlaps( void *h, const char *read, int strand, int barcode, int skipRepeats, int32_t *out, double *sim, int cap )
{
	SeqSet *s = (SeqSet *)h ;
	char *r = strdup( read ) ;
	std::vector<struct _overlap> overlaps ;
	int n = s->GetOverlapsFromRead( r, strand, barcode, 0, skipRepeats != 0, overlaps ) ;
	for ( int i = 0 ; i < n && i < cap ; ++i )
	{
		out[8 * i] = overlaps[i].seqIdx ;
		out[8 * i + 1] = overlaps[i].readStart ;
		out[8 * i + 2] = overlaps[i].readEnd ;
		out[8 * i + 3] = overlaps[i].seqStart ;
		out[8 * i + 4] = overlaps[i].seqEnd ;
		out[8 * i + 5] = overlaps[i].strand ;
		out[8 * i + 6] = overlaps[i].matchCnt ;
		out[8 * i + 7] = overlaps[i].indelCnt ;
		sim[i] = overlaps[i].similarity ;
	}
	free( r ) ;
	return n ;
}

// AlignAlgo::GlobalAlignment_PosWeight (AlignAlgo.hpp:57).  align must hold lent+lenp+2 entries.
int t4ref_dp_pos_weight( const int32_t *tWeights, int lent, const char *p, int lenp, signed char *align )
{
	struct _posWeight *w = new struct _posWeight[lent + 1] ;
	for ( int i = 0 ; i < lent ; ++i )
		for ( int j = 0 ; j < 4 ; ++j )
			w[i].count[j] = tWeights[4 * i + j] ;
	char *pp = (char *)malloc( lenp + 1 ) ;
	memcpy( pp, p, lenp ) ;
	pp[lenp] = '\0' ;
	int ret = (int)AlignAlgo::GlobalAlignment_PosWeight( w, lent, pp, lenp, align ) ;
	free( pp ) ;
	delete[] w ;
	return ret ;
}

// Postings of one k-mer (KmerIndex::Search, KmerIndex.hpp:104), in the reference's list order.
int t4ref_index_lookup( void *h, uint64_t code, int barcode, int32_t *out, int cap )
{
	SeqSet *s = (SeqSet *)h ;
	KmerCode kc( s->kmerLength ) ;
	kc.SetCode( code ) ;
	SimpleVector<struct _indexInfo> &l = *s->seqIndex.Search( kc, barcode ) ;
	int n = l.Size() ;
	for ( int i = 0 ; i < n && i < cap ; ++i )
	{
		out[2 * i] = l[i].idx ;
		out[2 * i + 1] = l[i].offset ;
	}
	return n ;
}

// Total number of postings and an order-independent checksum of the whole index.
int64_t t4ref_index_checksum( void *h, uint64_t *checksum )
{
	SeqSet *s = (SeqSet *)h ;
	int64_t total = 0 ;
	uint64_t sum = 0 ;
	for ( int b = 0 ; b < KINDEX_HASH_MAX ; ++b )
	{
		for ( std::map<uint64_t, SimpleVector<struct _indexInfo> >::iterator it = s->seqIndex.index[b].begin() ;
			it != s->seqIndex.index[b].end() ; ++it )
		{
			int n = it->second.Size() ;
			total += n ;
			for ( int i = 0 ; i < n ; ++i )
			{
				uint64_t x = it->first * 0x9E3779B97F4A7C15ull ^ ( (uint64_t)(uint32_t)it->second[i].idx << 32 | (uint32_t)it->second[i].offset ) ;
				x ^= x >> 31 ; x *= 0xBF58476D1CE4E5B9ull ; x ^= x >> 29 ;
				sum += x ;
			}
		}
	}
	*checksum = sum ;
	return total ;
}

int t4ref_nomatch_gap_limit( void *h ) { return ((SeqSet *)h)->nomatchGapLimit ; }

// Merge oracle of SURVEY.md section 8e(2): the reference's own in-tree merge idiom (main.cpp:2288-2294) applied to the
// concatenation of the shard contig sets in the given (rank, stream) order:
//   merged.InputSeqSet( shard, false ) for every shard ; merged.ChangeKmerLength( 31 ) ; merged.RemoveRedundantSeq()
// (SeqSet.hpp:3108, 4624, 4965: a contig that is a <= 1-mismatch substring of another one is dropped).
// Returns a new SeqSet handle (free with t4ref_destroy); *removed receives the number of contigs dropped.
void *t4ref_merge_sets( void *const *shards, int n, int kmerLength, int *removed )
{
	SeqSet *m = new SeqSet( kmerLength ) ;
	for ( int i = 0 ; i < n ; ++i )
		m->InputSeqSet( *(SeqSet *)shards[i], false ) ;
	int before = 0 ;
	for ( size_t i = 0 ; i < m->seqs.size() ; ++i )
		if ( m->seqs[i].consensus != NULL )
			++before ;
	m->ChangeKmerLength( 31 ) ;
	int after = m->RemoveRedundantSeq() ;
	if ( removed )
		*removed = before - after ;
	return m ;
}

// Load one contig (consensus + posWeight + name) into a reference SeqSet: lets a test rebuild shard sets from the
// packed buffers of t4_streams_pack_contigs and feed them to t4ref_merge_sets.
int t4ref_input_contig( void *h, const char *name, const char *consensus, const int32_t *posWeight, int barcode, int numRead )
{
	SeqSet *s = (SeqSet *)h ;
	int len = strlen( consensus ) ;
	struct _seqWrapper ns ;
	ns.name = strdup( name ) ;
	ns.consensus = strdup( consensus ) ;
	ns.consensusLen = len ;
	ns.isRef = false ;
	ns.minLeftExtAnchor = ns.minRightExtAnchor = 0 ;
	ns.barcode = barcode ;
	ns.numRead = numRead ;
	ns.index = true ;
	ns.posWeightCompressed = false ;
	for ( int j = 0 ; j < 3 ; ++j )
		ns.info[j].a = ns.info[j].b = ns.info[j].c = 0 ;
	int idx = s->seqs.size() ;
	s->seqs.push_back( ns ) ;
	struct _seqWrapper &sw = s->seqs[idx] ;
	sw.posWeight.ExpandTo( len ) ;
	for ( int i = 0 ; i < len ; ++i )
		for ( int k = 0 ; k < 4 ; ++k )
			sw.posWeight[i].count[k] = posWeight[4 * i + k] ;
	KmerCode kmerCode( s->kmerLength ) ;
	s->seqIndex.BuildIndexFromRead( kmerCode, sw.consensus, len, idx, barcode ) ;
	return idx ;
}

// ---- restated driver loop over read descriptors ---------------------------
// Follows main.cpp:1583-1880 (main pass) and 1897-1940 (rescue pass).  The static
// per-read quantities (filter, gene prefix, strand, thresholds, anchoring) arrive
// precomputed in t4_read_desc; see include/trust4_b200.h.
static double RescueThreshold( int minCnt ) // main.cpp:1913-1923
{
	double t = 0.9 ;
	if ( minCnt >= 20 )
		t = 0.97 ;
	else if ( minCnt >= 2 )
		t = 0.95 ;
	return t ;
}

int t4ref_run_descs( void *h, const t4_run_cfg *cfg, const t4_read_desc *descs, int n, const char *pool,
	const char *const *names, int nNames, int32_t *retCodes, int8_t *strands, int32_t *rescueRet )
{
	SeqSet *s = (SeqSet *)h ;
	int i, j ;
	int assembledReadCnt = 0 ;
	int prevAddRet = -1 ;
	int indexKmerLength = s->kmerLength ;
	int changeKmerLengthThreshold = cfg->change_k_threshold ;
	std::vector<char> goodCandidate( n, 0 ) ;
	std::vector<int> info( n, -1 ) ;
	std::vector<int> rescue ;
	// cfg->release_barcodes: purge finished barcodes like main.cpp:1572-1581, 1846-1859 (ReleaseFinishedBarcodeSeq with
	// release_index = true, contigMinCov = cfg->contig_min_cov, early_stop = true); note that the driver only counts reads
	// with addRet >= 0 towards "finished" (the block sits inside `else if ( addRet >= 0 )`).
	std::map<int, int> barcodeTotalReadCount, barcodeReadCount ;
	if ( cfg->has_barcode && cfg->release_barcodes )
		for ( i = 0 ; i < n ; ++i )
			if ( descs[i].barcode != -1 )
				++barcodeTotalReadCount[ descs[i].barcode ] ;
	char *buf = (char *)malloc( T4_MAX_READ_LEN + 2 ) ;
	for ( i = 0 ; i < n ; ++i )
	{
		const t4_read_desc &d = descs[i] ;
		int addRet = -1 ;
		memcpy( buf, pool + d.seq_off, d.len ) ;
		buf[d.len] = '\0' ;
		if ( !( d.flags & T4_RD_DUP ) )
		{
			int strand = 0 ;
			if ( d.flags & T4_RD_FILTERED )
				addRet = -1 ;
			else
			{
				char name[5] ;
				memcpy( name, d.gene4, 4 ) ;
				name[4] = '\0' ;
				strand = d.strand_in ;
				addRet = s->AddRead( buf, name, strand, d.barcode, d.min_kmer_count, cfg->repetitive != 0, d.sim_threshold ) ;
				if ( addRet < 0 )
				{
					if ( d.flags & T4_RD_NOVEL_ON_FAIL )
						addRet = s->InputNovelRead( names[d.name_id], buf, d.novel_strand, d.barcode ) ;
					else if ( d.flags & T4_RD_MOTIF_FORCED )
					{
						if ( s->HasMotif( buf, d.novel_strand ) )
							addRet = s->InputNovelRead( "Novel", buf, d.novel_strand, d.barcode ) ;
					}
					else if ( goodCandidate[i] )
					{
						int ms = -strands[ info[i] ] ;
						if ( ms != 0 && ( d.flags & T4_RD_MOTIF ) )
							addRet = s->InputNovelRead( "Novel", buf, ms, d.barcode ) ;
					}
				}
				strands[i] = strand ;
			}
			if ( d.flags & T4_RD_FILTERED )
				strands[i] = 0 ; // sortedReads[i].strand keeps its rough-annotation value; not observable downstream
		}
		else
		{
			if ( prevAddRet != -1 && prevAddRet != -3 )
				addRet = s->RepeatAddRead( buf ) ;
			else if ( prevAddRet == -3 )
				addRet = -3 ;
			strands[i] = i > 0 ? strands[i - 1] : 0 ;
		}

		if ( addRet == -2 )
			rescue.push_back( i ) ;
		else if ( addRet >= 0 )
		{
			++assembledReadCnt ;
			if ( d.mate_idx > i )
			{
				bool good = false ;
				if ( strands[i] == 1 && ( d.flags & T4_RD_GOOD_PLUS ) )
					good = true ;
				if ( strands[i] == -1 && ( d.flags & T4_RD_GOOD_MINUS ) )
					good = true ;
				if ( good && !goodCandidate[ d.mate_idx ] )
				{
					int tag = d.mate_idx ;
					const t4_read_desc &m = descs[tag] ;
					for ( j = tag - 1 ; j > 0 && j >= m.eq_lo ; --j )
					{
						goodCandidate[j] = 1 ;
						info[j] = i ;
					}
					for ( j = tag + 1 ; j < n && j < m.eq_hi ; ++j )
					{
						goodCandidate[j] = 1 ;
						info[j] = i ;
					}
				}
				if ( good )
				{
					goodCandidate[ d.mate_idx ] = 1 ;
					info[ d.mate_idx ] = i ;
				}
			}
		}
		if ( addRet >= 0 && cfg->has_barcode && cfg->release_barcodes && d.barcode != -1 )
		{
			int barcode = d.barcode ;
			++barcodeReadCount[barcode] ;
			if ( barcodeReadCount[barcode] >= barcodeTotalReadCount[barcode] )
			{
				std::map<int, int> finishedBarcodes ;
				finishedBarcodes[barcode] = barcodeTotalReadCount[barcode] ;
				s->ReleaseFinishedBarcodeSeq( finishedBarcodes, true, cfg->contig_min_cov, true ) ;
			}
		}
		retCodes[i] = addRet ;

		if ( assembledReadCnt > 0 && cfg->update_consensus_every > 0 && assembledReadCnt % cfg->update_consensus_every == 0
			&& !cfg->has_barcode )
			s->UpdateAllConsensus() ;
		prevAddRet = addRet ;
		if ( changeKmerLengthThreshold > 0 && s->Size() > changeKmerLengthThreshold && indexKmerLength < 16 && !cfg->has_barcode )
		{
			changeKmerLengthThreshold *= 4 ;
			indexKmerLength += 2 ;
			s->ChangeKmerLength( indexKmerLength ) ;
		}
	}
	if ( cfg->final_update )
		s->UpdateAllConsensus() ;

	if ( rescueRet != NULL )
		for ( i = 0 ; i < n ; ++i )
			rescueRet[i] = INT_MIN ;
	if ( cfg->do_rescue && cfg->first_read_len <= 200 )
	{
		int cnt = rescue.size() ;
		for ( i = 0 ; i < cnt ; ++i )
		{
			const t4_read_desc &d = descs[ rescue[i] ] ;
			memcpy( buf, pool + d.seq_off, d.len ) ;
			buf[d.len] = '\0' ;
			char name[2] = "" ;
			int strand = 0 ;
			int addRet = s->AddRead( buf, name, strand, d.barcode, 1, cfg->repetitive != 0, RescueThreshold( d.min_cnt ) ) ;
			strands[ rescue[i] ] = strand ;
			if ( rescueRet != NULL )
				rescueRet[ rescue[i] ] = addRet ;
		}
		if ( cfg->final_update )
			s->UpdateAllConsensus() ;
	}
	free( buf ) ;
	return assembledReadCnt ;
}

// ---- AssignRead pass --------------------------------------
// `SeqSet extendedSeq( k ) ; extendedSeq.InputSeqSet( seqSet, false )` (main.cpp:2047-2048, SeqSet.hpp:3108).
void *t4ref_input_seqset( void *src, int kmerLength )
{
	SeqSet *m = new SeqSet( kmerLength ) ;
	m->InputSeqSet( *(SeqSet *)src, false ) ;
	return m ;
}

// int SeqSet::AssignRead( read, strand, barcode, assign ), SeqSet.hpp:4632.  out[8] = {seqIdx, readStart, readEnd,
// seqStart, seqEnd, strand, matchCnt, 0} (the fields ExtendOverlap assigns, SeqSet.hpp:1248-1256) and *sim, written
// only when a contig was found; returns AssignRead's value.
int t4ref_assign_read( void *h, const char *read, int strand, int barcode, int32_t *out, double *sim )
{
	SeqSet *s = (SeqSet *)h ;
	struct _overlap assign ;
	memset( &assign, 0, sizeof( assign ) ) ;
	char *r = strdup( read ) ;
	int ret = s->AssignRead( r, strand, barcode, assign ) ;
	free( r ) ;
	if ( ret >= 0 )
	{
		out[0] = assign.seqIdx ; out[1] = assign.readStart ; out[2] = assign.readEnd ; out[3] = assign.seqStart ;
		out[4] = assign.seqEnd ; out[5] = assign.strand ; out[6] = assign.matchCnt ; out[7] = 0 ;
		*sim = assign.similarity ;
	}
	return ret ;
}

// The driver's AssignRead pass (main.cpp:2047-2118) over the stage-1 set `h`, restated on read descriptors:
//   extendedSeq( k ).InputSeqSet( seqSet, false ) ; SetNovelSeqSimilarity( 0.95 ) ;
//   for the assembled reads in the driver's order (list[], main.cpp:1779, 1933: main pass ascending, then the rescued
//   ones) AssignRead( read, sortedReads[].strand, barcode ) unless the read string equals its predecessor's in the list
//   (then the previous result is kept, main.cpp:2078-2081) ; SetNovelSeqSimilarity( 0.9 ) ; RecomputePosWeight.
// assign[8 * i], sim[i] describe list element i (seqIdx -1: not assigned; the reference leaves the other fields stale
// then and nothing reads them).  Returns the extended set (t4ref_destroy).
void *t4ref_assign_pass( void *h, int kmerLength, const t4_read_desc *descs, const char *pool, const int32_t *list, int nList,
	const int8_t *strands, int recompute, int32_t *assignOut, double *simOut )
{
	SeqSet *ext = new SeqSet( kmerLength ) ;
	ext->InputSeqSet( *(SeqSet *)h, false ) ;
	ext->SetNovelSeqSimilarity( 0.95 ) ;
	std::vector<struct _assignRead> reads( nList ) ;
	struct _overlap assign ;
	memset( &assign, 0, sizeof( assign ) ) ;
	assign.seqIdx = -1 ;
	for ( int i = 0 ; i < nList ; ++i )
	{
		const t4_read_desc &d = descs[ list[i] ] ;
		struct _assignRead &nr = reads[i] ;
		nr.id = NULL ;
		nr.read = (char *)malloc( d.len + 1 ) ;
		memcpy( nr.read, pool + d.seq_off, d.len ) ;
		nr.read[d.len] = '\0' ;
		nr.barcode = d.barcode ;
		nr.umi = -1 ;
		nr.info = list[i] ;
		nr.overlap.seqIdx = -1 ;
		nr.overlap.strand = strands[ list[i] ] ;
	}
	for ( int i = 0 ; i < nList ; ++i )
	{
		if ( i == 0 || strcmp( reads[i].read, reads[i - 1].read ) )
			ext->AssignRead( reads[i].read, reads[i].overlap.strand, reads[i].barcode, assign ) ;
		reads[i].overlap = assign ;
		int32_t *o = assignOut + 8 * i ;
		o[0] = assign.seqIdx ; o[1] = assign.readStart ; o[2] = assign.readEnd ; o[3] = assign.seqStart ;
		o[4] = assign.seqEnd ; o[5] = assign.strand ; o[6] = assign.matchCnt ; o[7] = 0 ;
		simOut[i] = assign.similarity ;
	}
	ext->SetNovelSeqSimilarity( 0.9 ) ;
	if ( recompute )
		ext->RecomputePosWeight( reads ) ;
	for ( int i = 0 ; i < nList ; ++i )
		free( reads[i].read ) ;
	return ext ;
}

// ---- k-mer counting ------------------------------------------
// KmerCount( k ): AddCount of every read, then GetCountStatsAndTrim( read, qual, ... ) of every read (main.cpp:404-440,
// 981-1010; qualPool == NULL: the qual == NULL call).  newLen[i] = strlen( read ) afterwards.
int t4ref_kmer_count_stats( const char *pool, const char *qualPool, const uint64_t *seqOff, const int32_t *len, int64_t n, int k,
	int32_t *minCnt, int32_t *medianCnt, float *avgCnt, int32_t *newLen )
{
	KmerCount kc( k ) ;
	int maxLen = 0 ;
	std::vector<char> buf, qbuf ;
	for ( int64_t i = 0 ; i < n ; ++i )
	{
		buf.assign( pool + seqOff[i], pool + seqOff[i] + len[i] ) ;
		buf.push_back( '\0' ) ;
		kc.AddCount( buf.data() ) ;
		if ( len[i] > maxLen )
			maxLen = len[i] ;
	}
	kc.SetBuffer( maxLen + 1 ) ;
	for ( int64_t i = 0 ; i < n ; ++i )
	{
		buf.assign( pool + seqOff[i], pool + seqOff[i] + len[i] ) ;
		buf.push_back( '\0' ) ;
		if ( qualPool )
		{
			qbuf.assign( qualPool + seqOff[i], qualPool + seqOff[i] + len[i] ) ;
			qbuf.push_back( '\0' ) ;
		}
		int a = 0, b = 0 ;
		float c = 0 ;
		kc.GetCountStatsAndTrim( buf.data(), qualPool ? qbuf.data() : NULL, a, b, c ) ;
		minCnt[i] = a ; medianCnt[i] = b ; avgCnt[i] = c ;
		if ( newLen )
			newLen[i] = (int)strlen( buf.data() ) ;
	}
	return 0 ;
}

// ---- stage-0 candidate extraction ----------------------------
// `SeqSet refSet( k ) ; refSet.InputRefFa( file )` (FastqExtractor.cpp:313-318)
void *t4ref_refset_create( const char *fasta, int k )
{
	SeqSet *s = new SeqSet( k ) ;
	s->InputRefFa( (char *)fasta ) ;
	return s ;
}
const char *t4ref_seq_name( void *h, int i ) { return ((SeqSet *)h)->seqs[i].name ; }
int t4ref_seq_count( void *h ) { return (int)((SeqSet *)h)->seqs.size() ; }
void t4ref_set_radius( void *h, int r ) { ((SeqSet *)h)->SetRadius( r ) ; }
// int SeqSet::HasHitInSet( char *read, int mode ), SeqSet.hpp:3144
int t4ref_has_hit_in_set( void *h, const char *read, int mode )
{
	char *r = strdup( read ) ;
	int ret = ((SeqSet *)h)->HasHitInSet( r, mode ) ;
	free( r ) ;
	return ret ;
}
// bool IsLowComplexity( char *seq ), FastqExtractor.cpp:106-127 -- a function of the extractor's main file, which cannot
// be included (it defines main); restated line by line.
int t4ref_is_low_complexity( const char *seq )
{
	int cnt[5] = {0, 0, 0, 0, 0} ;
	int i ;
	for ( i = 0 ; seq[i] ; ++i )
	{
		if ( seq[i] == 'N' )
			++cnt[4] ;
		else
			++cnt[ nucToNum[ seq[i] - 'A' ] ] ;
	}
	if ( cnt[0] >= i / 2 || cnt[1] >= i / 2 || cnt[2] >= i / 2 || cnt[3] >= i / 2 || cnt[4] >= i / 10 )
		return 1 ;
	int lowCnt = 0 ;
	for ( i = 0 ; i < 4 ; ++i )
		if ( cnt[i] <= 2 )
			++lowCnt ;
	if ( lowCnt >= 2 )
		return 1 ;
	return 0 ;
}

// SeqSet::LongestIncreasingSubsequence (SeqSet.hpp:342) on explicit (a, b) pairs sorted by b
int t4ref_lis( void *h, const int32_t *a, const int32_t *b, int n, int32_t *outA, int32_t *outB )
{
	SimpleVector<struct _pair> hits, lis ;
	for ( int i = 0 ; i < n ; ++i )
	{
		struct _pair p ;
		p.a = a[i] ; p.b = b[i] ;
		hits.PushBack( p ) ;
	}
	int ret = ((SeqSet *)h)->LongestIncreasingSubsequence( hits, lis ) ;
	for ( int i = 0 ; i < ret ; ++i )
	{
		outA[i] = lis[i].a ;
		outB[i] = lis[i].b ;
	}
	return ret ;
}

// int SeqSet::AnnotateRead( read, detailLevel 0, geneOverlap, NULL, NULL ), SeqSet.hpp:6016: out[t][8] for t = V, D, J, C,
// sim[t]; fresh `struct _overlap geneOverlap[4]` per call (the driver reuses one array, main.cpp:1086; an unset entry only
// has seqIdx = -1 and strand = 1 defined).
int t4ref_annotate_read( void *h, const char *read, int32_t *out, double *sim )
{
	struct _overlap geneOverlap[4] ;
	char *r = strdup( read ) ;
	int ret = ((SeqSet *)h)->AnnotateRead( r, 0, geneOverlap, NULL, NULL ) ;
	free( r ) ;
	for ( int t = 0 ; t < 4 ; ++t )
	{
		int32_t *o = out + 8 * t ;
		o[0] = geneOverlap[t].seqIdx ; o[1] = geneOverlap[t].readStart ; o[2] = geneOverlap[t].readEnd ; o[3] = geneOverlap[t].seqStart ;
		o[4] = geneOverlap[t].seqEnd ; o[5] = geneOverlap[t].strand ; o[6] = geneOverlap[t].matchCnt ; o[7] = geneOverlap[t].indelCnt ;
		sim[t] = geneOverlap[t].similarity ;
	}
	return ret ;
}

// std::sort( sortedReads ) of the stage-1 driver (main.cpp:1078).  `struct _sortRead` lives in main.cpp (which defines main
// and cannot be included), so its operator< (main.cpp:103-125) is restated here line by line on the fields it reads.
struct T4RefSortRead
{
	const char *id ;
	const char *read ;
	int minCnt, medianCnt ;
	float avgCnt ;
	int len ;
	int64_t idx ;
	bool operator<( const T4RefSortRead &b ) const
	{
		if ( minCnt != b.minCnt )
			return minCnt > b.minCnt ;
		else if ( medianCnt != b.medianCnt )
			return medianCnt > b.medianCnt ;
		else if ( avgCnt != b.avgCnt )
			return avgCnt > b.avgCnt ;
		else if ( len != b.len )
			return len > b.len ;
		else
		{
			int tmp = strcmp( read, b.read ) ;
			if ( tmp != 0 )
				return tmp < 0 ;
			else
				return strcmp( id, b.id ) < 0 ;
		}
	}
} ;

int t4ref_sort_reads( const char *const *reads, const char *const *ids, const int32_t *minCnt, const int32_t *medianCnt, const float *avgCnt,
	int64_t n, int64_t *order )
{
	std::vector<T4RefSortRead> v( n ) ;
	for ( int64_t i = 0 ; i < n ; ++i )
	{
		v[i].id = ids[i] ; v[i].read = reads[i] ; v[i].minCnt = minCnt[i] ; v[i].medianCnt = medianCnt[i] ; v[i].avgCnt = avgCnt[i] ;
		v[i].len = (int)strlen( reads[i] ) ; v[i].idx = i ;
	}
	std::sort( v.begin(), v.end() ) ;
	for ( int64_t i = 0 ; i < n ; ++i )
		order[i] = v[i].idx ;
	return 0 ;
}

// AlignAlgo::IsMateOverlap (AlignAlgo.hpp:1027); offset / bestMatchCnt start at -1 (the function only assigns offset when it
// finds a candidate)
int t4ref_is_mate_overlap( const char *fr, int flen, const char *sr, int slen, int minOverlap, int checkTandem, int32_t *offset, int32_t *bestMatchCnt )
{
	int off = -1, best = -1 ;
	int ret = AlignAlgo::IsMateOverlap( (char *)fr, flen, (char *)sr, slen, minOverlap, off, best, checkTandem != 0 ) ;
	*offset = off ;
	*bestMatchCnt = best ;
	return ret ;
}

} // extern "C"
