#!/usr/bin/env bash
# Builds the DROP-IN PROGRAM: the reference's own main / Simulator / XML loader / output code, compiled from
# /root/reference where it lies into THIS directory's _build/obj (the product does not share a build directory
# with the oracle), linked with
#   * this directory's Acceleration.cpp, RungeKutta4.cpp, RungeKuttaFehlberg78.cpp, DormandPrince.cpp INSTEAD of the
#     reference's four hot-path translation units (those four are never compiled here),
#   * Calculate.cpp overriding three members of Calculate (Integrals, PotentialEnergy, Energy),
#   * SavePhases.cpp overriding the one member BinaryFileAdapter::SavePhases,
#   * SimulatorHooks.cpp putting hooks in front of Simulator::BodyListToBodyData, CheckEvent and ShortestPeriod
#     (in COPIES of the reference's objects those symbols are renamed, so the originals stay callable),
#   * libsolaris_b200.so.
#   -> solaris_b200/host/_build/solaris_b200_dropin     (git-ignored; travels to the GPU box)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${SOLARIS_REF:-/root/reference}"
OUT="$HERE/_build"
OBJ="$OUT/obj"
if [ ! -d "$REF/Solaris" ]; then
  echo "build_dropin.sh: $REF/Solaris not present - keeping prebuilt drop-in (if any)" >&2
  exit 0
fi
mkdir -p "$OBJ"
# same flags as the oracle build: baseline x86-64, no FMA contraction, MSVC abs() semantics
ABSFIX="$HERE/absfix.h"
CXXFLAGS="-std=gnu++11 -O2 -w -fpermissive -fkeep-inline-functions -fPIC -ffp-contract=off -include cstring -include $ABSFIX -I$REF/Solaris -I$HERE"
pids=()
for f in "$REF"/Solaris/*.cpp; do
  b="$(basename "$f" .cpp)"
  case "$b" in Acceleration|RungeKutta4|RungeKuttaFehlberg78|DormandPrince) continue ;; esac
  o="$OBJ/$b.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
    g++ $CXXFLAGS -c "$f" -o "$o" &
    pids+=($!)
    if [ ${#pids[@]} -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
MINE="sol_bridge Acceleration RungeKutta4 RungeKuttaFehlberg78 DormandPrince Calculate SavePhases SimulatorHooks"
for f in $MINE; do
  g++ $CXXFLAGS -c "$HERE/$f.cpp" -o "$OUT/$f.o"
done
KEEP=$(ls "$OBJ"/*.o | grep -v -E '/(Calculate|BinaryFileAdapter|Simulator)\.o$')
# every member of BinaryFileAdapter stays the reference's; SavePhases is renamed in a copy of the object so that
# SavePhases.cpp can define the member and still call the original for the cases it does not handle
objcopy --redefine-sym _ZN17BinaryFileAdapter10SavePhasesEdiPdPiNS_10OutputTypeE=solb200_reference_SavePhases \
    "$OBJ/BinaryFileAdapter.o" "$OUT/BinaryFileAdapter_renamed.o"
# same for Calculate: Integrals / PotentialEnergy / Energy come from Calculate.cpp here (device), the O(n) members
# (TotalMass, PhaseOfBC, ...) stay the reference's
objcopy --redefine-sym _ZN9Calculate9IntegralsEP8BodyData=solb200_reference_Calculate_Integrals \
        --redefine-sym _ZN9Calculate15PotentialEnergyEP8BodyDataRd=solb200_reference_Calculate_PotentialEnergy \
        --redefine-sym _ZN9Calculate6EnergyEP8BodyDataRd=solb200_reference_Calculate_Energy \
    "$OBJ/Calculate.o" "$OUT/Calculate_renamed.o"
# Simulator: every member stays the reference's; BodyListToBodyData and CheckEvent get a hook in front
# (SimulatorHooks.cpp) that hands the event thresholds of Settings to the bridge and skips the host scan when the
# device found nothing to do.  Both are called from INSIDE Simulator.o (Integrate, DecisionMaking), so renaming the
# symbol would rename those references too: instead the reference's definitions are made WEAK (the hooks' strong
# definitions win at link time, for the intra-object calls as well - the object is compiled -fPIC, so they go
# through the symbol) and each gets a second, global name at the same address for the hooks to call.
alias_of() {  # <object> <mangled name> -> ".text:0x<offset>" of the symbol
  nm "$1" | awk -v s="$2" '$3 == s && ($2 == "T" || $2 == "W") { print ".text:0x" $1 }'
}
A_BL=$(alias_of "$OBJ/Simulator.o" _ZN9Simulator18BodyListToBodyDataEv)
A_CE=$(alias_of "$OBJ/Simulator.o" _ZN9Simulator10CheckEventEd)
A_SP=$(alias_of "$OBJ/Simulator.o" _ZN9Simulator14ShortestPeriodEv)
[ -n "$A_BL" ] && [ -n "$A_CE" ] && [ -n "$A_SP" ] || { echo "build_dropin.sh: Simulator symbols not found" >&2; exit 1; }
objcopy --weaken-symbol=_ZN9Simulator18BodyListToBodyDataEv --weaken-symbol=_ZN9Simulator10CheckEventEd \
        --weaken-symbol=_ZN9Simulator14ShortestPeriodEv \
        --add-symbol solb200_reference_Simulator_BodyListToBodyData=$A_BL,global,function \
        --add-symbol solb200_reference_Simulator_CheckEvent=$A_CE,global,function \
        --add-symbol solb200_reference_Simulator_ShortestPeriod=$A_SP,global,function \
    "$OBJ/Simulator.o" "$OUT/Simulator_renamed.o"
OBJS=""
for f in $MINE; do OBJS="$OBJS $OUT/$f.o"; done
g++ -o "$OUT/solaris_b200_dropin" $KEEP $OBJS "$OUT"/BinaryFileAdapter_renamed.o "$OUT"/Calculate_renamed.o \
    "$OUT"/Simulator_renamed.o -L"$ROOT/solaris_b200" -lsolaris_b200 -lpthread \
    -Wl,-rpath,'$ORIGIN/../..' -Wl,-rpath,/usr/local/cuda/lib64
echo "built $OUT/solaris_b200_dropin"
