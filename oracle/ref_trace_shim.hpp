// TEST INFRASTRUCTURE (oracle) -- not part of the product; nothing in trust4_b200/ may use it.
//
// Force-included (g++ -include) in front of the UNMODIFIED reference main.cpp to
// record every call the stage-1 driver makes on the novel-contig SeqSet
// (/root/reference/main.cpp:642 `SeqSet seqSet`), i.e. the traffic over the
// drop-in boundary of SURVEY.md section 8b.  The reference sources are compiled
// where they lie; nothing is copied.
//
// Mechanism: SeqSet.hpp is included here first (its include guard then makes
// main.cpp's own #include a no-op), a subclass hides the boundary methods with
// logging wrappers, and `#define SeqSet` makes main.cpp instantiate the subclass.
// Only the first-constructed instance (seqSet, main.cpp:642) is logged.
//
// Output: text lines on $T4_TRACE_OUT (tab separated), see tests/trace_format.md.
#ifndef T4_REF_TRACE_SHIM
#define T4_REF_TRACE_SHIM
#include <stdarg.h>
#include <time.h>
#include <assert.h>
#include <map>
#include "SeqSet.hpp"

static FILE *t4_trace_fp()
{
	static FILE *fp = NULL ;
	static bool init = false ;
	if ( !init )
	{
		init = true ;
		const char *p = getenv( "T4_TRACE_OUT" ) ;
		if ( p != NULL )
			fp = fopen( p, "w" ) ;
	}
	return fp ;
}

static int t4_trace_instances = 0 ;

class T4TraceSeqSet : public SeqSet
{
	int inst ;
	bool Log() { return inst == 0 && t4_trace_fp() != NULL ; }
public:
	T4TraceSeqSet( int kl ) : SeqSet( kl )
	{
		inst = t4_trace_instances++ ;
		if ( Log() ) fprintf( t4_trace_fp(), "C\t%d\n", kl ) ;
	}
	int AddRead( char *read, char *geneName, int &strand, int barcode, int minKmerCount, bool repetitiveData, double similarityThreshold )
	{
		int strandIn = strand ;
		int ret = SeqSet::AddRead( read, geneName, strand, barcode, minKmerCount, repetitiveData, similarityThreshold ) ;
		if ( Log() )
			fprintf( t4_trace_fp(), "A\t%s\t%s\t%d\t%d\t%d\t%d\t%.17g\t%d\t%d\n", read, geneName[0] ? geneName : ".", strandIn,
				barcode, minKmerCount, repetitiveData ? 1 : 0, similarityThreshold, ret, strand ) ;
		return ret ;
	}
	int RepeatAddRead( char *read )
	{
		int ret = SeqSet::RepeatAddRead( read ) ;
		if ( Log() ) fprintf( t4_trace_fp(), "R\t%s\t%d\n", read, ret ) ;
		return ret ;
	}
	int InputNovelRead( const char *id, char *read, int strand, int barcode )
	{
		int ret = SeqSet::InputNovelRead( id, read, strand, barcode ) ;
		if ( Log() ) fprintf( t4_trace_fp(), "N\t%s\t%s\t%d\t%d\t%d\n", id, read, strand, barcode, ret ) ;
		return ret ;
	}
	int HasMotif( char *read, int strand )
	{
		int ret = SeqSet::HasMotif( read, strand ) ;
		if ( Log() ) fprintf( t4_trace_fp(), "M\t%s\t%d\t%d\n", read, strand, ret ) ;
		return ret ;
	}
	void UpdateAllConsensus()
	{
		SeqSet::UpdateAllConsensus() ;
		if ( Log() ) fprintf( t4_trace_fp(), "U\n" ) ;
	}
	void ChangeKmerLength( int kl )
	{
		SeqSet::ChangeKmerLength( kl ) ;
		if ( Log() ) fprintf( t4_trace_fp(), "K\t%d\n", kl ) ;
	}
	int SetHitLenRequired( int l )
	{
		if ( Log() ) fprintf( t4_trace_fp(), "H\t%d\n", l ) ;
		return SeqSet::SetHitLenRequired( l ) ;
	}
	void SetIsLongSeqSet( bool in )
	{
		if ( Log() ) fprintf( t4_trace_fp(), "L\t%d\n", in ? 1 : 0 ) ;
		SeqSet::SetIsLongSeqSet( in ) ;
	}
	void SetConsiderBarcodeInIndexHash( bool s )
	{
		if ( Log() ) fprintf( t4_trace_fp(), "B\t%d\n", s ? 1 : 0 ) ;
		SeqSet::SetConsiderBarcodeInIndexHash( s ) ;
	}
	void ReleaseFinishedBarcodeSeq( std::map<int, int> &finished, bool releaseIndex, int minCov, bool earlyStop )
	{
		if ( Log() )
		{
			fprintf( t4_trace_fp(), "F\t%d\t%d\t%d", releaseIndex ? 1 : 0, minCov, earlyStop ? 1 : 0 ) ;
			for ( std::map<int, int>::iterator it = finished.begin() ; it != finished.end() ; ++it )
				fprintf( t4_trace_fp(), "\t%d:%d", it->first, it->second ) ;
			fprintf( t4_trace_fp(), "\n" ) ;
		}
		SeqSet::ReleaseFinishedBarcodeSeq( finished, releaseIndex, minCov, earlyStop ) ;
	}
	void ReleaseShallowContigs( int minCov )
	{
		if ( Log() ) fprintf( t4_trace_fp(), "S\t%d\n", minCov ) ;
		SeqSet::ReleaseShallowContigs( minCov ) ;
	}
	void Output( FILE *fp, std::vector<std::string> *barcodeIntToStr = NULL )
	{
		if ( Log() ) { fprintf( t4_trace_fp(), "O\n" ) ; fflush( t4_trace_fp() ) ; }
		SeqSet::Output( fp, barcodeIntToStr ) ;
	}
} ;

#define SeqSet T4TraceSeqSet
#endif
