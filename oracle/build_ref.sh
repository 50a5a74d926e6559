#!/usr/bin/env bash
# Build the UNMODIFIED reference (suliaron/solaris, Solaris/*.cpp) as the parity oracle.
#
#   oracle/_ref/solaris_ref          the reference program (main in Solaris.cpp)
#   oracle/_ref/libsolaris_ref.a     everything except main, for in-process harnesses
#   oracle/_ref/libref_harness.so    C-ABI harness (oracle/ref_harness.cpp) over the reference's own
#                                    Acceleration / RungeKutta4 / RungeKuttaFehlberg78 / DormandPrince
#
# Sources are compiled where they lie under $SOLARIS_REF (default /root/reference); nothing is copied
# into this repository and nothing is written outside oracle/_ref/.  Flags follow SURVEY.md §8(c):
# baseline x86-64 (no FMA contraction), -O2, the abs() fix of oracle/absfix.h.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SOLARIS_REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/Solaris" ]; then
  echo "build_ref.sh: $REF/Solaris not present - keeping prebuilt oracle/_ref (if any)" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
CXXFLAGS="-std=gnu++11 -O2 -w -fpermissive -fkeep-inline-functions -fPIC -ffp-contract=off -include cstring -include $HERE/absfix.h -I$REF/Solaris"
pids=()
for f in "$REF"/Solaris/*.cpp; do
  o="$OUT/obj/$(basename "$f" .cpp).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
    g++ $CXXFLAGS -c "$f" -o "$o" &
    pids+=($!)
    if [ ${#pids[@]} -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
g++ -o "$OUT/solaris_ref" "$OUT"/obj/*.o
rm -f "$OUT/libsolaris_ref.a"
ar rcs "$OUT/libsolaris_ref.a" $(ls "$OUT"/obj/*.o | grep -v '/Solaris.o')
if [ -f "$HERE/ref_harness.cpp" ]; then
  g++ $CXXFLAGS -shared -o "$OUT/libref_harness.so" "$HERE/ref_harness.cpp" \
      -Wl,--whole-archive "$OUT/libsolaris_ref.a" -Wl,--no-whole-archive -lpthread
fi
echo "oracle/_ref built from $REF"
