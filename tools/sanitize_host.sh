#!/bin/bash
# Host-side sanitizer pass (build container, no GPU): the three C++ front-ends are built against the oracle shim with
# AddressSanitizer + UndefinedBehaviorSanitizer and again with ThreadSanitizer, then driven by the fuzzers (which
# compare every run with the reference programs, so a sanitizer report shows up as a mismatch: reports change the exit
# status).  handle_segv/handle_abort are off because several recorded behaviours ARE a SIGSEGV / SIGABRT.
#   tools/sanitize_host.sh [runs-per-fuzzer, default 100]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
RUNS=${1:-100}
H=$ROOT/acfcalculator_b200/csrc/host
S=$ROOT/tests/host/abi_shim_oracle.cpp
CXX=$([ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
make -s -C "$ROOT/oracle" >/dev/null
COMMON="$S $H/options.cpp $H/readers.cpp $H/fast_readers.cpp $H/gle.cpp $H/exp_cr.cpp"
LINK="-pthread -L$ROOT/oracle -loracle -Wl,-rpath,$ROOT/oracle -fopenmp"
for kind in asan tsan; do
    out=$(mktemp -d)
    if [ $kind = asan ]; then
        FL="-O1 -g -std=c++14 -fsanitize=address,undefined -fno-omit-frame-pointer -fno-sanitize-recover=undefined"
        export ASAN_OPTIONS=handle_segv=0:handle_abort=0:detect_leaks=0
    else
        FL="-O1 -g -std=c++14 -fsanitize=thread"
        export TSAN_OPTIONS="handle_segv=0 handle_abort=0"
    fi
    $CXX $FL -o $out/ACFcalculator_oracle $H/acfcalculator_main.cpp $COMMON $LINK
    $CXX $FL -o $out/viscosity_oracle $H/viscosity_main.cpp $COMMON $LINK
    $CXX $FL -o $out/traj2diffusion_oracle $H/traj2diffusion_main.cpp $H/t2d_readers.cpp $COMMON $LINK
    echo "== $kind"
    FUZZ_OURS=$out/ACFcalculator_oracle python "$ROOT/tools/fuzz_readers.py" 101 $RUNS | tail -n 1
    FUZZ_OURS=$out/ACFcalculator_oracle python "$ROOT/tools/fuzz_options.py" 101 $((RUNS * 5)) | tail -n 1
    FUZZ_OURS=$out/ACFcalculator_oracle python "$ROOT/tools/fuzz_gle.py" 101 $((RUNS / 4 + 1)) | tail -n 1
    FUZZ_OURS=$out/viscosity_oracle python "$ROOT/tools/fuzz_viscosity.py" 101 $((RUNS / 2 + 1)) | tail -n 1
    FUZZ_OURS=$out/traj2diffusion_oracle python "$ROOT/tools/fuzz_t2d.py" 101 $RUNS | tail -n 1
    rm -rf $out
done
