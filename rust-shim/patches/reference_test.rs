// Add to the `tests` module of summa-dev/halo2-experiments src/circuits/merkle_sum_tree.rs, next to test_full_prover
// (/root/reference/src/circuits/merkle_sum_tree.rs:345-358).  Same circuit instance, same public inputs; OsRng replaced
// by a seeded XorShiftRng inside RecordingRng so that the run is reproducible and every draw is captured.
#[test]
fn test_dump_vectors_for_b200zk() {
    use b200zk_shim::RecordingRng;
    use halo2_proofs::{halo2curves::bn256::{Bn256, G1Affine}, plonk::{create_proof, keygen_pk, keygen_vk},
        poly::kzg::{commitment::{KZGCommitmentScheme, ParamsKZG}, multiopen::ProverSHPLONK},
        transcript::{Blake2bWrite, Challenge255, TranscriptWriterBuffer}};
    use rand_core::SeedableRng;
    use rand_xorshift::XorShiftRng;
    const SEED: [u8; 16] = [0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5];

    let (circuit, public_input) = instantiate_circuit(500);          // test_full_prover's instance: [10, 100, root, 500], k = 9
    let k = 9;
    let mut srs_rng = RecordingRng::new(XorShiftRng::from_seed(SEED));
    let params = ParamsKZG::<Bn256>::setup(k, &mut srs_rng);        // srs_rng.bytes = srs_secret_wide (64 bytes)
    let vk = keygen_vk(&params, &circuit).expect("vk");
    let pk = keygen_pk(&params, vk, &circuit).expect("pk");
    let mut rng = RecordingRng::new(XorShiftRng::from_seed(SEED));  // a fresh stream for the prover, as tests/ here assume
    let mut transcript = Blake2bWrite::<_, G1Affine, Challenge255<_>>::init(vec![]);
    create_proof::<KZGCommitmentScheme<Bn256>, ProverSHPLONK<'_, Bn256>, Challenge255<G1Affine>, _, _, _>(
        &params, &pk, &[circuit], &[&[&public_input]], &mut rng, &mut transcript).expect("prover should not fail");
    let _proof = transcript.finalize();                              // the dump hook inside create_proof wrote B200ZK_DUMP_DIR
}
