#!/usr/bin/env python3
"""Generate the batch-route driver: the reference main.cpp with ONE inserted line.

    python make_batch_main.py <reference main.cpp> <output .cpp>

The line `T4_BATCH_PREPARE() ;` goes directly in front of the stage-1 AddRead loop (main.cpp:1583, the first
`for ( i = 0 ; i < readCnt ; ++i )` after `int prevAddRet = -1 ;`).  The macro is defined by t4_seqset_adapter.hpp
(force-included); everything else of main.cpp is used as it is.  The output is a build artefact (integration/_build/,
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
    out = lines[:loop] + ["\tT4_BATCH_PREPARE() ; // trust4_b200 batch route (integration/t4_seqset_adapter.hpp)",
                          '#line %d "%s"' % (loop + 1, src)] + lines[loop:]
    open(dst, "w").write("\n".join(out))
    print("make_batch_main: inserted T4_BATCH_PREPARE() before line %d of %s" % (loop + 1, src))


if __name__ == "__main__":
    main()
