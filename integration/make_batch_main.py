#!/usr/bin/env python3
"""Generate the batch-route driver: the reference main.cpp with TWO inserted lines (one per offloaded pass).

    python make_batch_main.py <reference main.cpp> <output .cpp>

The line `T4_BATCH_PREPARE() ;` goes directly in front of the stage-1 AddRead loop (main.cpp:1583, the first
`for ( i = 0 ; i < readCnt ; ++i )` after `int prevAddRet = -1 ;`).  The line `T4_BATCH_ASSIGN()` (an `if ( !... )` without
a body) goes directly in front of the AssignRead loop (main.cpp:2075, the `if ( threadCnt <= 1 ) ... else ...` statement
after `extendedSeq.SetNovelSeqSimilarity( 0.95 ) ;`), which becomes its fall-back branch.  Both macros are defined by
t4_seqset_adapter.hpp (force-included); everything else of main.cpp is used as it is.  The output is a build artefact (integration/_build/,
git-ignored) -- no reference source is kept in this repository."""
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = open(src).read().split("\n")
    start = [i for i, l in enumerate(lines) if l.strip() == "int prevAddRet = -1 ;"]
    if len(start) != 1:
        sys.exit("make_batch_main: anchor 'int prevAddRet = -1 ;' not found exactly once")
    loop = None
    for i in range(start[0], len(lines)):
        if lines[i] == "\tfor ( i = 0 ; i < readCnt ; ++i )":
            loop = i
            break
    if loop is None or loop - start[0] > 80:
        sys.exit("make_batch_main: the AddRead loop header was not found after the anchor")
    sim = [i for i, l in enumerate(lines) if l.strip() == "extendedSeq.SetNovelSeqSimilarity( 0.95 ) ;"]
    if len(sim) != 1 or lines[sim[0] + 1] != "\tif ( threadCnt <= 1 )":
        sys.exit("make_batch_main: the AssignRead loop (`if ( threadCnt <= 1 )` after SetNovelSeqSimilarity( 0.95 )) was not found")
    asg = sim[0] + 1
    out = (lines[:loop] + ["\tT4_BATCH_PREPARE() ; // trust4_b200 batch route (integration/t4_seqset_adapter.hpp)",
                           '#line %d "%s"' % (loop + 1, src)] + lines[loop:asg]
           + ["\tT4_BATCH_ASSIGN() // trust4_b200 batch route: the statement below is the fall-back branch",
              '#line %d "%s"' % (asg + 1, src)] + lines[asg:])
    open(dst, "w").write("\n".join(out))
    print("make_batch_main: inserted T4_BATCH_PREPARE() before line %d and T4_BATCH_ASSIGN() before line %d of %s" % (loop + 1, asg + 1, src))


if __name__ == "__main__":
    main()
