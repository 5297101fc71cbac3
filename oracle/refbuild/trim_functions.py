#!/usr/bin/env python3
"""Build helper (TEST INFRASTRUCTURE): writes a copy of a reference source file WITHOUT the bodies of the named member
functions - what a maintainer does by hand when the shim/*.cc files take those functions over. The copy goes
to a temporary directory that the Makefile deletes after compiling; nothing of it is kept in the repository.

usage: trim_functions.py SRC DST 'first-line regex' ['first-line regex' ...]
Each regex must match exactly one line: the first line of a function definition at namespace scope. The definition is removed
through its matching closing brace."""
import re
import sys


def main():
    src, dst, pats = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = open(src, encoding="utf-8", errors="replace").read().split("\n")
    drop = set()
    for pat in pats:
        hits = [i for i, l in enumerate(lines) if re.search(pat, l)]
        if len(hits) != 1:
            sys.exit("trim_functions: %r matches %d lines in %s" % (pat, len(hits), src))
        i = hits[0]
        depth, seen, j = 0, False, i
        while True:
            for ch in lines[j]:
                if ch == "{":
                    depth += 1
                    seen = True
                elif ch == "}":
                    depth -= 1
            if seen and depth == 0:
                break
            j += 1
        drop.update(range(i, j + 1))
    with open(dst, "w", encoding="utf-8") as f:
        f.write("\n".join("" if k in drop else l for k, l in enumerate(lines)))  # line numbers of the rest are preserved


if __name__ == "__main__":
    main()
