#!/bin/bash
# Pins motivate-b200's oracle on a REAL run of the reference.  Needs a Rust toolchain (not in the build image, so
# this has never been run there) and a checkout of 0xr0bert/Motivate with rust/patch/dump_state.rs applied as
# rust/patch/README.md §B says.
#     tools/make_ref_fixtures.sh /path/to/Motivate [people]
# Result: tests/golden/ref_c1_tiefree.npz; `pytest tests/test_reference_fixtures.py` then compares the oracle (and,
# with -m gpu, the CUDA path) with the Rust run day by day.
set -euo pipefail
REF=${1:?path to the patched Motivate checkout}
PEOPLE=${2:-5000}                      # small: the fixture stores every agent's state for every weekday
REPO=$(cd "$(dirname "$0")/.." && pwd)
WORK=$(mktemp -d)
cp "$REPO/rust/patch/dump_state.rs" "$REF/src/dump_state.rs"
( cd "$REF" && cargo build --release )
# the shipped scenario, three fixes applied, ties removed, scaled to $PEOPLE, 1 simulation, 2 years (intervention on day 365)
python - "$WORK" "$PEOPLE" <<PY
import sys; sys.path.insert(0, "$REPO")
from motivate_b200 import c1
c1.materialise(sys.argv[1], people=int(sys.argv[2]), simulations=1, total_years=2, tie_free=True)
PY
( cd "$WORK" && MOTIVATE_DUMP_DIR="$WORK/dump" "$REF/target/release/motivate" --generate )
python "$REPO/tools/ref_dump_to_npz.py" "$WORK/dump/1" "$REPO/tests/golden/ref_c1_tiefree.npz"
cp "$WORK/output/output_1.csv" "$REPO/tests/golden/ref_c1_tiefree.csv"
echo "fixture written; run: python -m pytest tests/test_reference_fixtures.py -q"
