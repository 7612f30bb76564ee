#!/bin/bash
# Race check of the kernel sources: the host emulation (lanes = host threads, one barrier per lane mask) built with
# ThreadSanitizer and driven through a few small cases.  A report names the two source lines of a shared-memory access pair
# that no barrier orders -- on the GPU such a pair is only safe while the lanes of a warp happen to run in lock step.
#   tests/host_emulation/tsan.sh [build|nmpc|nmpc8|nmpc_vns|dtc|sim|sim_vns|sim_wb|sim_cycle|soft|soft_idx|est ...]      (CPU only, a few minutes; needs gcc's libtsan)
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd); CSRC=$ROOT/model-predictive-control-tuning_b200/csrc
g++ -O1 -g -std=c++20 -fPIC -shared -pthread -fsanitize=thread -x c++ -o /tmp/libmpcemu_tsan.so $HERE/emu.cpp $HERE/emu_nmpc.cpp $HERE/emu_dtc.cpp $HERE/emu_build.cpp \
    $CSRC/mpc_tables.cpp $CSRC/mpc_dtc_tables.cpp -lm
for w in ${@:-build nmpc nmpc8 nmpc_vns dtc sim sim_vns sim_wb sim_cycle soft soft_idx est}; do
  echo "== $w"
  LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" \
    python $HERE/tsan_cases.py $w 2>&1 | grep -E "WARNING: ThreadSanitizer|    #[01] .*\.cuh|^ok" | awk '!seen[$0]++' | head -40
done
