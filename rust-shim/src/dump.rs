//! Writes one prove job together with stock halo2's proof as a directory of little-endian binaries + manifest.json
//! (layout: rust-shim/README.md).  `tools/import_rust_vectors.py` turns it into tests/golden/rust_<name>.npz, and
//! tests/test_rust_vectors.py then asserts that the oracle and the GPU path reproduce halo2's bytes.
use crate::cs_blob;
use halo2_proofs::halo2curves::bn256::{Fr, G1Affine, G2Affine};
use halo2_proofs::plonk::ConstraintSystem;
use std::fs;
use std::io::Write;
use std::path::Path;

fn raw<T>(v: &[T]) -> &[u8] {
    // Fr / G1Affine / G2Affine are plain arrays of u64 Montgomery limbs: dumped as they sit in memory
    unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u8, std::mem::size_of_val(v)) }
}
fn put(dir: &Path, name: &str, bytes: &[u8]) {
    fs::File::create(dir.join(name)).and_then(|mut f| f.write_all(bytes)).expect("dump write");
}
fn flat<T: Copy>(cols: &[Vec<T>]) -> Vec<T> {
    cols.iter().flat_map(|c| c.iter().copied()).collect()
}

pub struct ProveJob<'a> {
    pub k: u32,
    pub srs_secret_wide: &'a [u8],      // the 64 bytes ParamsKZG::setup drew
    pub cs: &'a ConstraintSystem<Fr>,   // vk.cs(): selectors already compressed
    pub fixed: &'a [Vec<Fr>],           // pk.fixed_values
    pub map_col: &'a [u32],             // permutation Assembly mapping, P x n
    pub map_row: &'a [u32],
    pub advice: &'a [Vec<Fr>],          // after batch_invert_assigned, before blinding
    pub instances: &'a [Vec<Fr>],
    pub rng_wide: &'a [u8],             // RecordingRng::bytes of create_proof
    pub transcript_repr: Fr,
    pub fixed_commitments: &'a [G1Affine],
    pub sigma_commitments: &'a [G1Affine],
    pub g2: G2Affine,
    pub s_g2: G2Affine,
    pub proof: &'a [u8],
    pub params_bytes: Option<&'a [u8]>, // ParamsKZG::write
}

pub fn dump(dir: &Path, job: &ProveJob) {
    fs::create_dir_all(dir).expect("dump dir");
    put(dir, "srs_secret_wide.bin", job.srs_secret_wide);
    let blob = cs_blob(job.cs, job.k);
    put(dir, "cs_blob.u32", raw(&blob));
    put(dir, "fixed.fr", raw(&flat(job.fixed)));
    put(dir, "map_col.u32", raw(job.map_col));
    put(dir, "map_row.u32", raw(job.map_row));
    put(dir, "advice.fr", raw(&flat(job.advice)));
    put(dir, "instances.fr", raw(&flat(job.instances)));
    put(dir, "rng_wide.bin", job.rng_wide);
    put(dir, "transcript_repr.fr", raw(&[job.transcript_repr]));
    put(dir, "fixed_commitments.g1", raw(job.fixed_commitments));
    put(dir, "sigma_commitments.g1", raw(job.sigma_commitments));
    put(dir, "g2.g2", raw(&[job.g2]));
    put(dir, "s_g2.g2", raw(&[job.s_g2]));
    put(dir, "proof.bin", job.proof);
    if let Some(p) = job.params_bytes {
        put(dir, "params.bin", p);
    }
    let manifest = serde_json::json!({
        "format": "b200zk-rust-vectors-1",
        "halo2_proofs": "privacy-scaling-explorations/halo2 tag v2023_02_02",
        "k": job.k,
        "num_fixed": job.fixed.len(), "num_advice": job.advice.len(), "num_permutation": job.sigma_commitments.len(),
        "instance_lens": job.instances.iter().map(|c| c.len()).collect::<Vec<_>>(),
        "rng_draws": job.rng_wide.len() / 64,
        "proof_bytes": job.proof.len(),
        "has_params_bytes": job.params_bytes.is_some(),
    });
    put(dir, "manifest.json", serde_json::to_string_pretty(&manifest).unwrap().as_bytes());
}
