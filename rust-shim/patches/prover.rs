// halo2_proofs/src/plonk/prover.rs (tag v2023_02_02)
//
// (A) create_proof through the device: witness synthesis stays exactly as upstream; everything after it is one call.
pub fn create_proof_gpu<R: RngCore, T: TranscriptWrite<G1Affine, Challenge255<G1Affine>>, ConcreteCircuit: Circuit<Fr>>(
    params: &ParamsKZG<Bn256>, pk: &ProvingKey<G1Affine>, circuits: &[ConcreteCircuit], instances: &[&[&[Fr]]], mut rng: R, transcript: &mut T,
) -> Result<(), Error> {
    assert_eq!(circuits.len(), 1, "b200zk_create_proof proves one circuit instance per call");
    // 1. WitnessCollection + batch_invert_assigned, unchanged (single phase: the reference's circuits use no challenges)
    let advice: Vec<Vec<Fr>> = synthesize_advice(params, pk, &circuits[0], instances[0])?;
    // 2. every Fr::random(rng) create_proof would make, in its order, as the raw 64 bytes each call consumes
    let dev = pk.dev.get_or_init(|| b200zk_shim::Pk::create(&params.dev, pk.vk.cs(), &pk.fixed_values_vec(), &pk.map_col(), &pk.map_row()));
    let mut wide = vec![0u8; 64 * dev.rng_draws()];
    rng.fill_bytes(&mut wide);
    // 3. the library runs the Blake2b transcript itself, starting from vk.transcript_repr
    let proof = dev.create_proof(&advice, instances[0], &wide, &pk.vk.transcript_repr).map_err(|_| Error::ConstraintSystemFailure)?;
    transcript.append_raw(&proof);          // small addition to TranscriptWrite: the bytes are final
    Ok(())
}

// (B) dump hook: paste at the end of STOCK create_proof (feature "b200zk-dump").  `advice_unblinded` must be captured
// before the blinding rows are written (clone `advice.advice_polys` right after batch_invert_assigned); `rng` is the
// caller's RecordingRng, `proof_so_far` the transcript bytes after the multiopen proof.
#[cfg(feature = "b200zk-dump")]
if let Ok(dir) = std::env::var("B200ZK_DUMP_DIR") {
    b200zk_shim::dump::dump(std::path::Path::new(&dir), &b200zk_shim::dump::ProveJob {
        k: params.k(),
        srs_secret_wide: &B200ZK_SRS_SECRET.lock().unwrap(),          // stored by ParamsKZG::setup under the same feature
        cs: pk.vk.cs(),
        fixed: &pk.fixed_values.iter().map(|p| p.values.clone()).collect::<Vec<_>>(),
        map_col: &pk.permutation_mapping_cols(), map_row: &pk.permutation_mapping_rows(),
        advice: &advice_unblinded,
        instances: &instances[0].iter().map(|c| c.to_vec()).collect::<Vec<_>>(),
        rng_wide: &rng.bytes,
        transcript_repr: pk.vk.transcript_repr,
        fixed_commitments: &pk.vk.fixed_commitments, sigma_commitments: pk.vk.permutation.commitments(),
        g2: params.g2(), s_g2: params.s_g2(),
        proof: &proof_so_far,
        params_bytes: None,
    });
}
