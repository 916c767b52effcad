// gen_goldens.rs -- golden-vector generator to be run ONCE on a machine with a Rust toolchain, inside a
// checkout of Safari77/rupphash (the reference).  It calls the reference's own functions on the four JPEG
// fixtures the reference ships and dumps what rupphash_b200's oracle and CUDA path must reproduce.
// The build image of rupphash_b200 has no cargo/rustc, so this file is a specification that has never been
// compiled there; tools/regen_with_rust.md says where to put it and how to feed the result to the tests.
//
// Output: one JSON object per line on stdout:
//   {"file": "...", "w": W, "h": H, "decoder": "image",            // pixels come from the `image` crate here
//    "rgb_sha256": "...",                                          // so that a PIL decode can be told apart
//    "pdq_hash": "64 hex", "pdq_quality": f32-bits, "pdq_coeffs": [256 x f32-bits],
//    "pdq_dihedral": ["64 hex" x 8],
//    "resized": {"w": .., "h": .., "luma_sha256": "..."},          // fast_image_resize Box output (if any)
//    "phash": "16 hex", "phash_dihedral": ["16 hex" x 8], "phash_rotation_invariant": "16 hex"}
//
// Place as src/bin/gen_goldens.rs and add to Cargo.toml:
//   [[bin]] name = "gen_goldens"  path = "src/bin/gen_goldens.rs"
// The reference's modules are private to its binaries; include them by path exactly as phash_test does
// (phash_test.rs:1-6): `#[path = "../pdqhash.rs"] mod pdqhash;` etc.
#[path = "../pdqhash.rs"]
mod pdqhash;
#[path = "../phash.rs"]
mod phash;

use sha2::{Digest, Sha256}; // add `sha2 = "0.10"` to [dependencies] (or drop the two digests)

fn hex(bytes: &[u8]) -> String {
    bytes.iter().map(|b| format!("{:02x}", b)).collect()
}

fn main() {
    let files = [
        "tests/bench.jpg",
        "tests/Kaanapali_beach_sunrise_on_Maui_Hawaii.720p.jpg",
        "tests/Prophecy_Has_Been_Fulfilled_1.jpg",
        "tests/Prophecy_Has_Been_Fulfilled_2.jpg",
    ];
    let hasher = phash::DctPhash::new();
    for f in files {
        let img = image::open(f).expect("decode");
        let rgb = img.to_rgb8();
        let (w, h) = (rgb.width(), rgb.height());
        // hot path 1: pdqhash::generate_pdq_features (pdqhash.rs:166-196) + to_hash / dihedral (:59-87)
        let (features, quality) = pdqhash::generate_pdq_features(&img).expect("image is at least 5x5");
        let hash = features.to_hash();
        let dihedral = features.generate_dihedral_hashes();
        // the pre-downsample alone (pdqhash.rs:172-191): expose `to_luma601`, `calculate_target_dimensions` and
        // `resize_luma_fast` as pub(crate) for this binary, or copy the three calls here
        let resized = {
            let luma = pdqhash::to_luma601(&img);
            if w > 512 || h > 512 {
                let (nw, nh) = pdqhash::calculate_target_dimensions(w, h, 512);
                let r = pdqhash::resize_luma_fast(&luma, nw, nh).expect("resize");
                format!("{{\"w\": {}, \"h\": {}, \"luma_sha256\": \"{}\"}}", nw, nh, hex(&Sha256::digest(r.as_raw())))
            } else {
                "null".to_string()
            }
        };
        // pHash (phash.rs:48-83, :137-255)
        let ph = hasher.hash_image(&img);
        let ph_dihedral = phash::generate_dihedral_hashes(ph);
        let ph_inv = phash::calculate_rotation_invariant_hash(ph);
        println!(
            "{{\"file\": \"{}\", \"w\": {}, \"h\": {}, \"decoder\": \"image\", \"rgb_sha256\": \"{}\", \"pdq_hash\": \"{}\", \
             \"pdq_quality\": {}, \"pdq_coeffs\": [{}], \"pdq_dihedral\": [{}], \"resized\": {}, \"phash\": \"{:016x}\", \
             \"phash_dihedral\": [{}], \"phash_rotation_invariant\": \"{:016x}\"}}",
            f, w, h, hex(&Sha256::digest(rgb.as_raw())), hex(&hash), quality.to_bits(),
            features.coefficients.iter().map(|c| c.to_bits().to_string()).collect::<Vec<_>>().join(", "),
            dihedral.iter().map(|d| format!("\"{}\"", hex(d))).collect::<Vec<_>>().join(", "),
            resized, ph,
            ph_dihedral.iter().map(|d| format!("\"{:016x}\"", d)).collect::<Vec<_>>().join(", "), ph_inv
        );
        // the decoded pixels themselves, so that the oracle can be fed EXACTLY what the reference hashed
        // (zune-jpeg / image differ from libjpeg-turbo by +-1 LSB): tests/golden/<name>.rgb
        std::fs::write(format!("{}.rgb", f), rgb.as_raw()).expect("write rgb");
    }
}
