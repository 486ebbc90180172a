#!/bin/bash
# Pins this repository's frozen vectors -- and with them the oracle and the CUDA path -- to the real plonky2 v0.2.0 prover.
# Needs what this image lacks: cargo with the toolchain of /root/reference/rust-toolchain (nightly-2024-02-22) and network
# (or vendored) access to plonky2 @ 7445ec911b0c5a1d94062ef2d5cae4ec08ee9a3f (Cargo.lock:4847-4850).  starkyx @ 1644af58
# (Cargo.lock:7232-7234) is needed in addition for the two cubic-extension gates and the AIR evaluator, not for this pin.
#
#   tools/pin_from_rust.sh            -> tests/golden/rust_pin.npz (cap, leaves, Merkle paths, hash_pad(&[]), Rayon Melem/s)
#   python -m pytest tests/test_rust_pin.py        consumes it: CPU oracle always, the GPU path with -m gpu
set -euo pipefail
cd "$(dirname "$0")/.."
command -v cargo >/dev/null || { echo "pin_from_rust.sh: cargo not found -- run this where the VectorX toolchain is installed" >&2; exit 2; }
python bench/rust/make_input.py
( cd bench/rust
  cargo run --release -- golden_values.bin golden_out.bin | tee golden_run.json
  cargo run --release -- config1_values.bin config1_out.bin 5 | tee config1_run.json
  cargo run --release -- --hash-pad > hash_pad.txt )
python bench/rust/check_output.py bench/rust/golden_out.bin --write tests/golden/rust_pin.npz \
       --hash-pad bench/rust/hash_pad.txt --timing bench/rust/config1_run.json
echo "wrote tests/golden/rust_pin.npz; now: python -m pytest tests/test_rust_pin.py (and -m gpu on a B200)"
