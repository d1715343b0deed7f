// Reference-side binding of libtrust4_b200.so: what a TRUST4 maintainer adds to run stage 1 (`trust4`, main.cpp)
// on the GPU without touching main.cpp.  Force-included in front of the UNMODIFIED reference main.cpp
// (integration/Makefile: g++ -include t4_seqset_adapter.hpp /root/reference/main.cpp -ltrust4_b200), it
//   * derives from the reference SeqSet, so everything outside the hot path (reference gene set, rough annotation,
//     AssignRead / mate extension, RemoveRedundantSeq) keeps running the reference's own CPU code;
//   * routes the novel-contig set of the stage-1 driver (`SeqSet seqSet`, main.cpp:642 -- the first SeqSet constructed)
//     through the C ABI of include/trust4_b200.h: AddRead, RepeatAddRead, InputNovelRead, UpdateAllConsensus,
//     ChangeKmerLength, Size, SetHitLenRequired, Output (SeqSet.hpp:3426, 4477, 3028, 4525, 4624, 2591, 2601, 10939),
//     ReleaseFinishedBarcodeSeq, ReleaseShallowContigs, InputNovelFa (SeqSet.hpp:10815, 10928, 2986);
//   * hands the contigs back to the CPU object after the assembly (Output is the driver's first use of them,
//     main.cpp:1957-1975) so that `extendedSeq.InputSeqSet( seqSet, false )` (main.cpp:2049) sees them.
// The resulting binary is a drop-in `trust4`: same flags, same _raw.out / _final.out / _assembled_reads.fa
// (tests/test_dropin_cli.py compares them byte for byte with the stock binary).
#ifndef T4_SEQSET_ADAPTER_HPP
#define T4_SEQSET_ADAPTER_HPP

#ifdef T4_ADAPTER_EMU /* test build against the CPU emulation library (tests/emu): same ABI, t4emu_ prefix */
#define t4_seqset_create t4emu_seqset_create
#define t4_seqset_destroy t4emu_seqset_destroy
#define t4_seqset_add_read t4emu_seqset_add_read
#define t4_seqset_repeat_add_read t4emu_seqset_repeat_add_read
#define t4_seqset_input_novel_read t4emu_seqset_input_novel_read
#define t4_seqset_update_all_consensus t4emu_seqset_update_all_consensus
#define t4_seqset_change_kmer_length t4emu_seqset_change_kmer_length
#define t4_seqset_size t4emu_seqset_size
#define t4_seqset_set_hit_len_required t4emu_seqset_set_hit_len_required
#define t4_seqset_set_is_long t4emu_seqset_set_is_long
#define t4_seqset_set_consider_barcode_in_hash t4emu_seqset_set_consider_barcode_in_hash
#define t4_seqset_output t4emu_seqset_output
#define t4_seqset_get_contig t4emu_seqset_get_contig
#define t4_seqset_contig_flags t4emu_seqset_contig_flags
#define t4_seqset_release_finished_barcode t4emu_seqset_release_finished_barcode
#define t4_seqset_release_shallow_contigs t4emu_seqset_release_shallow_contigs
#define t4_seqset_input_novel_fa t4emu_seqset_input_novel_fa
#define t4_last_error t4emu_last_error
#define t4_init t4emu_init
#endif

#include <stdarg.h>
#include <time.h>
#include <assert.h>
#include <map>
#include <string>
#include <vector>
#define private public /* the adapter refills SeqSet::seqs from the device */
#include "SeqSet.hpp"
#undef private
#include "trust4_b200.h"

static int t4_adapter_instances = 0 ;

class T4GpuSeqSet : public SeqSet
{
	t4_seqset *h ;
	bool gpu ;
	void Die() { fprintf( stderr, "trust4_b200: %s\n", t4_last_error() ) ; exit( 1 ) ; }
	int Check( int r ) { if ( r < T4_E_BASE ) Die() ; return r ; }

	// Copy the device contigs into the CPU object (slot numbers preserved; released slots stay NULL).
	void SyncToHost()
	{
		int n = Check( t4_seqset_size( h ) ) ;
		for ( size_t i = 0 ; i < seqs.size() ; ++i )
		{
			if ( seqs[i].consensus ) free( seqs[i].consensus ) ;
			if ( seqs[i].name ) free( seqs[i].name ) ;
		}
		seqs.clear() ;
		std::vector<char> cons, name( 4096 ) ;
		std::vector<int32_t> pw ;
		std::map<int, int> purgedBarcodes ;
		for ( int i = 0 ; i < n ; ++i )
		{
			struct _seqWrapper ns ;
			ns.name = ns.consensus = NULL ;
			ns.consensusLen = 0 ;
			ns.isRef = false ;
			ns.minLeftExtAnchor = ns.minRightExtAnchor = 0 ;
			ns.barcode = -1 ;
			ns.numRead = 0 ;
			ns.index = true ;
			ns.posWeightCompressed = false ;
			for ( int j = 0 ; j < 3 ; ++j )
				ns.info[j].a = ns.info[j].b = ns.info[j].c = 0 ;
			int len = t4_seqset_get_contig( h, i, NULL, 0, NULL, NULL, 0, NULL, NULL, NULL, NULL ) ;
			if ( len >= 0 )
			{
				cons.resize( len + 1 ) ;
				pw.resize( 4 * (size_t)len + 4 ) ;
				Check( t4_seqset_get_contig( h, i, cons.data(), len + 1, pw.data(), name.data(), (int)name.size(), &ns.barcode,
					&ns.numRead, &ns.minLeftExtAnchor, &ns.minRightExtAnchor ) ) ;
				ns.consensus = strdup( cons.data() ) ;
				ns.name = strdup( name.data() ) ;
				ns.consensusLen = len ;
				if ( Check( t4_seqset_contig_flags( h, i ) ) & T4_CONTIG_PURGED )
					purgedBarcodes[ns.barcode] = 1 ;
			}
			seqs.push_back( ns ) ;
			if ( len >= 0 )
			{
				struct _seqWrapper &sw = seqs[i] ;
				sw.posWeight.ExpandTo( len ) ;
				for ( int j = 0 ; j < len ; ++j )
					for ( int k = 0 ; k < 4 ; ++k )
						sw.posWeight[j].count[k] = pw[4 * j + k] ;
			}
		}
		// Contigs the device purged (ReleaseFinishedBarcodeSeq) keep their full posWeight columns there; the reference
		// compresses or frees them and marks them un-indexed.  Re-apply exactly that storage change to the mirror with the
		// reference's own code (the index removal and UpdateConsensus inside it are no-ops here: the host index is empty
		// and the consensus is already final), so everything downstream of Output sees the reference's object state.
		if ( !purgedBarcodes.empty() )
			SeqSet::ReleaseFinishedBarcodeSeq( purgedBarcodes, true, 0, false ) ;
	}
public:
	T4GpuSeqSet( int kl ) : SeqSet( kl ), h( NULL )
	{
		gpu = ( t4_adapter_instances++ == 0 ) ;
		if ( gpu )
		{
			h = t4_seqset_create( kl ) ;
			if ( !h )
				Die() ;
		}
	}
	~T4GpuSeqSet() { if ( h ) t4_seqset_destroy( h ) ; }

	int AddRead( char *read, char *geneName, int &strand, int barcode, int minKmerCount, bool repetitiveData, double similarityThreshold )
	{
		if ( !gpu )
			return SeqSet::AddRead( read, geneName, strand, barcode, minKmerCount, repetitiveData, similarityThreshold ) ;
		return Check( t4_seqset_add_read( h, read, geneName, &strand, barcode, minKmerCount, repetitiveData, similarityThreshold ) ) ;
	}
	int RepeatAddRead( char *read ) { return gpu ? Check( t4_seqset_repeat_add_read( h, read ) ) : SeqSet::RepeatAddRead( read ) ; }
	int InputNovelRead( const char *id, char *read, int strand, int barcode )
	{
		return gpu ? Check( t4_seqset_input_novel_read( h, id, read, strand, barcode ) ) : SeqSet::InputNovelRead( id, read, strand, barcode ) ;
	}
	void UpdateAllConsensus() { if ( gpu ) Check( t4_seqset_update_all_consensus( h ) ) ; else SeqSet::UpdateAllConsensus() ; }
	void ChangeKmerLength( int kl )
	{
		if ( gpu )
			Check( t4_seqset_change_kmer_length( h, kl ) ) ;
		SeqSet::ChangeKmerLength( kl ) ; // keeps kmerLength of the CPU object in step (its seqs are empty until SyncToHost)
	}
	int Size() { return gpu ? Check( t4_seqset_size( h ) ) : SeqSet::Size() ; }
	int SetHitLenRequired( int l )
	{
		if ( gpu )
			Check( t4_seqset_set_hit_len_required( h, l ) ) ;
		return SeqSet::SetHitLenRequired( l ) ;
	}
	void SetIsLongSeqSet( bool in )
	{
		if ( gpu )
			Check( t4_seqset_set_is_long( h, in ) ) ;
		SeqSet::SetIsLongSeqSet( in ) ;
	}
	void SetConsiderBarcodeInIndexHash( bool s )
	{
		if ( gpu )
			Check( t4_seqset_set_consider_barcode_in_hash( h, s ) ) ;
		SeqSet::SetConsiderBarcodeInIndexHash( s ) ;
	}
	// main.cpp:1855 -- one finished barcode, removeFromIndex = true, earlyStop = true
	void ReleaseFinishedBarcodeSeq( std::map<int, int> barcodes, bool removeFromIndex, int contigMinCov, bool earlyStop )
	{
		if ( !gpu )
		{
			SeqSet::ReleaseFinishedBarcodeSeq( barcodes, removeFromIndex, contigMinCov, earlyStop ) ;
			return ;
		}
		if ( barcodes.size() != 1 || !removeFromIndex || !earlyStop )
		{
			fprintf( stderr, "trust4_b200: ReleaseFinishedBarcodeSeq is only supported as the stage-1 driver calls it\n" ) ;
			exit( 1 ) ;
		}
		Check( t4_seqset_release_finished_barcode( h, barcodes.begin()->first, contigMinCov ) ) ;
	}
	// main.cpp:1954 (--contigMinCov)
	void ReleaseShallowContigs( int minCov )
	{
		if ( gpu )
			Check( t4_seqset_release_shallow_contigs( h, minCov ) ) ;
		else
			SeqSet::ReleaseShallowContigs( minCov ) ;
	}
	// main.cpp:711 (--debug-ns)
	void InputNovelFa( char *filename )
	{
		if ( gpu )
			Check( t4_seqset_input_novel_fa( h, filename ) ) ;
		else
			SeqSet::InputNovelFa( filename ) ;
	}
	void Output( FILE *fp, std::vector<std::string> *barcodeIntToStr = NULL )
	{
		if ( !gpu )
		{
			SeqSet::Output( fp, barcodeIntToStr ) ;
			return ;
		}
		std::vector<const char *> names ;
		if ( barcodeIntToStr )
			for ( size_t i = 0 ; i < barcodeIntToStr->size() ; ++i )
				names.push_back( barcodeIntToStr->at( i ).c_str() ) ;
		Check( t4_seqset_output( h, fp, barcodeIntToStr ? names.data() : NULL, (int)names.size() ) ) ;
		SyncToHost() ; // from here on the driver only reads the contigs (main.cpp:2049 InputSeqSet)
	}
} ;

#define SeqSet T4GpuSeqSet
#endif
