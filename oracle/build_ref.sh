#!/usr/bin/env bash
# oracle/build_ref.sh — compile the UNMODIFIED reference solvers into oracle/_ref/.
# TEST INFRASTRUCTURE ONLY.  Sources are read where they lie (REFERENCE_DIR,
# default /root/reference); parameter variants are made by text substitution
# into a scratch copy under /tmp, never into this repo.
# Outputs go only to oracle/_ref/ (git-ignored; not gpurun-ignored).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${REFERENCE_DIR:-/root/reference}"
out="$here/_ref"
if [ ! -d "$ref" ]; then echo "build_ref: $ref absent, keeping prebuilt $out" >&2; exit 0; fi
mkdir -p "$out"
scratch="$(mktemp -d /tmp/pm_refsrc_XXXXXX)"
trap 'rm -rf "$scratch"' EXIT
# the image's $CXX (/opt/gcc wrapper) links libstdc++ statically, which breaks iostreams inside a dlopen()ed .so
CXX="${PM_HOST_CXX:-$([ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)}"
FLAGS="-std=c++17 -O2 -ffp-contract=off -shared -fPIC -w"

build() { # name case src
  "$CXX" $FLAGS -DREF_CASE="$2" -DREF_SRC="\"$3\"" "$here/ref_wrap.cpp" -o "$out/libref_$1.so"
  echo "built $out/libref_$1.so"
}
must_change() { # file_before file_after what
  if cmp -s "$1" "$2"; then echo "build_ref: substitution '$3' did not apply" >&2; exit 1; fi
}

# 1. verbatim defaults
build cavity_default  0 "$ref/cavity-01.cpp"
build channel_default 1 "$ref/channel-01.cpp"
build step_default    2 "$ref/backwards_step-01.cpp"

# 2. BASELINE cfg[0]: cavity Re=100, 128x128, dt=1e-3
sed -e 's/n_interior          = 63;/n_interior          = 128;/' \
    -e 's/reynolds_number  = 1000.0;/reynolds_number  = 100.0;/' \
    -e 's/time_step(cfl_number \* std::min(/time_step(1e-3 + 0.0 * std::min(/' \
    "$ref/cavity-01.cpp" > "$scratch/cavity_cfg0.cpp"
must_change "$ref/cavity-01.cpp" "$scratch/cavity_cfg0.cpp" cavity_cfg0
grep -q 'n_interior          = 128;' "$scratch/cavity_cfg0.cpp" && grep -q '1e-3 + 0.0' "$scratch/cavity_cfg0.cpp" && grep -q 'reynolds_number  = 100.0' "$scratch/cavity_cfg0.cpp"
build cavity_cfg0 0 "$scratch/cavity_cfg0.cpp"

# 3. BASELINE cfg[1]: channel Re=1000, 256x64, dt=5e-4
sed -e 's/NX_INT             = 93;/NX_INT             = 256;/' \
    -e 's/NY_INT             = 31;/NY_INT             = 64;/' \
    -e 's/REYNOLDS_NUMBER    = 100.0;/REYNOLDS_NUMBER    = 1000.0;/' \
    -e 's/time_step(CFL \* std::min(/time_step(5e-4 + 0.0 * std::min(/' \
    "$ref/channel-01.cpp" > "$scratch/channel_cfg1.cpp"
grep -q 'NX_INT             = 256;' "$scratch/channel_cfg1.cpp" && grep -q 'NY_INT             = 64;' "$scratch/channel_cfg1.cpp" && grep -q '5e-4 + 0.0' "$scratch/channel_cfg1.cpp" && grep -q 'REYNOLDS_NUMBER    = 1000.0' "$scratch/channel_cfg1.cpp"
build channel_cfg1 1 "$scratch/channel_cfg1.cpp"

# 4. benchmark shapes of SURVEY §8d for the CPU arm: cavity Re=1000, K=100 cap, 512^2 / 1024^2 / 2048^2
for n in 512 1024 2048; do
  sed -e "s/n_interior          = 63;/n_interior          = $n;/" \
      -e 's/max_sor_iterations  = 10000;/max_sor_iterations  = 100;/' \
      "$ref/cavity-01.cpp" > "$scratch/cavity_k100_$n.cpp"
  grep -q "n_interior          = $n;" "$scratch/cavity_k100_$n.cpp" && grep -q 'max_sor_iterations  = 100;' "$scratch/cavity_k100_$n.cpp"
  build "cavity_k100_$n" 0 "$scratch/cavity_k100_$n.cpp"
done

# 5. small fixed-K variants used by parity tests (cap 50 iterations so a step costs milliseconds)
sed -e 's/n_interior          = 63;/n_interior          = 32;/' \
    -e 's/max_sor_iterations  = 10000;/max_sor_iterations  = 50;/' \
    "$ref/cavity-01.cpp" > "$scratch/cavity_k50_32.cpp"
grep -q 'n_interior          = 32;' "$scratch/cavity_k50_32.cpp"
build cavity_k50_32 0 "$scratch/cavity_k50_32.cpp"
sed -e 's/MAX_SOR_ITERS      = 10000;/MAX_SOR_ITERS      = 50;/' "$ref/channel-01.cpp" > "$scratch/channel_k50.cpp"
grep -q 'MAX_SOR_ITERS      = 50;' "$scratch/channel_k50.cpp"
build channel_k50 1 "$scratch/channel_k50.cpp"
sed -e 's/MAX_SOR_ITERS      = 10000;/MAX_SOR_ITERS      = 50;/' "$ref/backwards_step-01.cpp" > "$scratch/step_k50.cpp"
grep -q 'MAX_SOR_ITERS      = 50;' "$scratch/step_k50.cpp"
build step_k50 2 "$scratch/step_k50.cpp"
