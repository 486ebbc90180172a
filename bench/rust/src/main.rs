//! Runs the reference's own commitment (plonky2 v0.2.0 `PolynomialBatch::from_values`, the function reached from
//! contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75) on a value matrix written by
//! `make_input.py`, writes the Merkle cap + a few leaves and Merkle paths for `check_output.py`, and prints the time of
//! the commit: the byte-level pin of the CUDA path against the real prover, and the published CPU number for BASELINE.md.
//!
//! usage: refcheck <values.bin> <out.bin> [repeats]
//!        refcheck --hash-pad                 prints PoseidonHash::hash_pad(&[]) (the domain-separator digest that
//!                                            CircuitBuilder::build hashes into circuit_digest) as 4 decimal u64
//! values.bin: u32 c, u32 log_n, u32 rate_bits, u32 cap_height, then c * 2^log_n u64 (column-major, little endian)
//! out.bin:    2^cap_height * 4 u64 cap, then for each probe index (0, 1, N/2, N-1): the leaf row (c u64) and its
//!             Merkle path ((log2 N - cap_height) * 4 u64)
use std::fs::File;
use std::io::{Read, Write};
use std::time::Instant;

use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::hash::poseidon::PoseidonHash;
use plonky2::plonk::config::Hasher;
use plonky2::plonk::config::PoseidonGoldilocksConfig as C;
use plonky2::util::timing::TimingTree;

const D: usize = 2;

fn read_u32(f: &mut File) -> anyhow::Result<u32> {
    let mut b = [0u8; 4];
    f.read_exact(&mut b)?;
    Ok(u32::from_le_bytes(b))
}

fn main() -> anyhow::Result<()> {
    let args: Vec<String> = std::env::args().collect();
    if args.len() == 2 && args[1] == "--hash-pad" {
        let h = <PoseidonHash as Hasher<F>>::hash_pad(&[]);
        let v: Vec<String> = h.elements.iter().map(|e| e.to_canonical_u64().to_string()).collect();
        println!("{}", v.join(" "));
        return Ok(());
    }
    anyhow::ensure!(args.len() >= 3, "usage: refcheck <values.bin> <out.bin> [repeats]");
    let repeats: usize = args.get(3).map(|s| s.parse()).transpose()?.unwrap_or(3);
    let mut f = File::open(&args[1])?;
    let (c, log_n) = (read_u32(&mut f)? as usize, read_u32(&mut f)? as usize);
    let (rate_bits, cap_height) = (read_u32(&mut f)? as usize, read_u32(&mut f)? as usize);
    let n = 1usize << log_n;
    let mut raw = vec![0u8; 8 * c * n];
    f.read_exact(&mut raw)?;
    let column = |j: usize| -> PolynomialValues<F> {
        let vals = (0..n)
            .map(|i| {
                let o = 8 * (j * n + i);
                F::from_canonical_u64(u64::from_le_bytes(raw[o..o + 8].try_into().unwrap()))
            })
            .collect();
        PolynomialValues::new(vals)
    };

    let mut best = f64::MAX;
    let mut batch = None;
    for _ in 0..repeats {
        let values: Vec<PolynomialValues<F>> = (0..c).map(column).collect();
        let mut timing = TimingTree::default();
        let t = Instant::now();
        // blinding = false, no FFT root table: exactly how the VectorX circuits and starkyx call it
        let b = PolynomialBatch::<F, C, D>::from_values(values, rate_bits, false, cap_height, &mut timing, None);
        best = best.min(t.elapsed().as_secs_f64());
        batch = Some(b);
    }
    let batch = batch.unwrap();
    let threads = std::thread::available_parallelism().map(|x| x.get()).unwrap_or(1);
    println!(
        "{{\"impl\": \"plonky2 v0.2.0 (Rayon)\", \"c\": {c}, \"log_n\": {log_n}, \"rate_bits\": {rate_bits}, \
         \"cap_height\": {cap_height}, \"seconds\": {best:.6}, \"melem_per_s\": {:.3}, \"threads\": {threads}}}",
        (c * n) as f64 / best / 1e6
    );

    let mut out = File::create(&args[2])?;
    let mut put = |x: F| out.write_all(&x.to_canonical_u64().to_le_bytes());
    for h in batch.merkle_tree.cap.0.iter() {
        for e in h.elements.iter() {
            put(*e)?;
        }
    }
    let big_n = n << rate_bits;
    for idx in [0usize, 1, big_n / 2, big_n - 1] {
        for e in batch.merkle_tree.get(idx).iter() {
            put(*e)?;
        }
        for h in batch.merkle_tree.prove(idx).siblings.iter() {
            for e in h.elements.iter() {
                put(*e)?;
            }
        }
    }
    Ok(())
}
