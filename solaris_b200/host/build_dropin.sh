#!/usr/bin/env bash
# Builds the DROP-IN PROGRAM: the reference's own main / Simulator / XML loader / output code (compiled
# from /root/reference where it lies, objects under oracle/_ref/obj) linked with THIS directory's
# Acceleration.cpp, RungeKutta4.cpp, RungeKuttaFehlberg78.cpp, DormandPrince.cpp instead of the reference's four
# translation units, Calculate.cpp overriding three members of Calculate (Integrals, PotentialEnergy, Energy), with SavePhases.cpp overriding the one member BinaryFileAdapter::SavePhases
# (in a COPY of the reference's object that symbol is renamed, so the original stays callable), and with libsolaris_b200.so.
#   -> solaris_b200/host/_build/solaris_b200_dropin     (git-ignored; travels to the GPU box)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${SOLARIS_REF:-/root/reference}"
OBJ="$ROOT/oracle/_ref/obj"
OUT="$HERE/_build"
if [ ! -d "$REF/Solaris" ]; then
  echo "build_dropin.sh: $REF/Solaris not present - keeping prebuilt drop-in (if any)" >&2
  exit 0
fi
[ -d "$OBJ" ] || "$ROOT/oracle/build_ref.sh" >/dev/null
mkdir -p "$OUT"
CXXFLAGS="-std=gnu++11 -O2 -w -fpermissive -fPIC -ffp-contract=off -include cstring -include $ROOT/oracle/absfix.h -I$REF/Solaris -I$HERE"
for f in sol_bridge Acceleration RungeKutta4 RungeKuttaFehlberg78 DormandPrince Calculate SavePhases; do
  g++ $CXXFLAGS -c "$HERE/$f.cpp" -o "$OUT/$f.o"
done
KEEP=$(ls "$OBJ"/*.o | grep -v -E '/(Acceleration|RungeKutta4|RungeKuttaFehlberg78|DormandPrince|Calculate|BinaryFileAdapter)\.o$')
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
g++ -o "$OUT/solaris_b200_dropin" $KEEP "$OUT"/sol_bridge.o "$OUT"/Acceleration.o "$OUT"/RungeKutta4.o \
    "$OUT"/RungeKuttaFehlberg78.o "$OUT"/DormandPrince.o "$OUT"/Calculate.o "$OUT"/SavePhases.o "$OUT"/BinaryFileAdapter_renamed.o "$OUT"/Calculate_renamed.o -L"$ROOT/solaris_b200" -lsolaris_b200 \
    -Wl,-rpath,'$ORIGIN/../..' -Wl,-rpath,/usr/local/cuda/lib64
echo "built $OUT/solaris_b200_dropin"
