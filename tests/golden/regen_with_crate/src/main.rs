//! Writes the compressed outputs behind tests/golden/known_answers.json, produced by the real lz4_flex crate,
//! as files `<out>/<name>.<mode>.lz4`; tests/golden/regen_with_crate/compare.py then hashes them against
//! known_answers.json (which holds outputs of this repo's C oracle).
//!
//!   cargo run --release -- <lz4_flex checkout>/benches <out dir>
//!
//! Modes (tests/golden/make_golden.py):
//!   block_api   lz4_flex::block::compress(input)
//!   frame_fresh the first block of a FrameEncoder frame (independent blocks): 5-byte hash, u32 table, offset 0
//!   frame_cont  the second block of a frame whose first block was other data: stream offset > 0
//!   frames      whole frames with the FrameInfo options listed in known_answers.json
use std::fs;
use std::io::Write;
use std::path::Path;

use lz4_flex::frame::{BlockMode, BlockSize, FrameEncoder, FrameInfo};

fn block_size_for(len: usize) -> Option<(BlockSize, usize)> {
    for (bs, n) in [
        (BlockSize::Max64KB, 64usize << 10),
        (BlockSize::Max256KB, 256 << 10),
        (BlockSize::Max1MB, 1 << 20),
        (BlockSize::Max4MB, 4 << 20),
        (BlockSize::Max8MB, 8 << 20),
    ] {
        if len <= n {
            return Some((bs, n));
        }
    }
    None
}

/// (BlockInfo word, payload) list of a frame without checksums / content size.
fn frame_blocks(frame: &[u8]) -> Vec<(u32, Vec<u8>)> {
    let mut pos = 7; // magic 4 + FLG + BD + HC
    let mut out = Vec::new();
    loop {
        let w = u32::from_le_bytes(frame[pos..pos + 4].try_into().unwrap());
        pos += 4;
        if w == 0 {
            break;
        }
        let len = (w & 0x7fff_ffff) as usize;
        out.push((w, frame[pos..pos + len].to_vec()));
        pos += len;
    }
    out
}

/// The n-th block (0 or 1) of a frame of independent blocks as the encoder compressed it; None if it was stored raw.
fn frame_block(input: &[u8], second: bool) -> Option<Vec<u8>> {
    let (bs, n) = block_size_for(input.len())?;
    let info = FrameInfo::new().block_size(bs).block_mode(BlockMode::Independent);
    let mut enc = FrameEncoder::with_frame_info(info, Vec::new());
    if second {
        // any full first block: its table entries lie below the second block's stream offset and never match
        let filler: Vec<u8> = (0..n).map(|i| (i * 131 % 251) as u8).collect();
        enc.write_all(&filler).unwrap();
        enc.flush().unwrap();
    }
    enc.write_all(input).unwrap();
    let frame = enc.finish().unwrap();
    let blocks = frame_blocks(&frame);
    let (w, payload) = blocks.into_iter().nth(if second { 1 } else { 0 })?;
    if w & 0x8000_0000 != 0 {
        None
    } else {
        Some(payload)
    }
}

fn whole_frame(input: &[u8], block_size_id: u32, flags: u32) -> Vec<u8> {
    let bs = match block_size_id {
        0 => BlockSize::Auto,
        4 => BlockSize::Max64KB,
        5 => BlockSize::Max256KB,
        6 => BlockSize::Max1MB,
        7 => BlockSize::Max4MB,
        _ => BlockSize::Max8MB,
    };
    let mut info = FrameInfo::new().block_size(bs).block_mode(BlockMode::Independent);
    info = info.block_checksums(flags & 1 != 0).content_checksum(flags & 2 != 0);
    if flags & 4 != 0 {
        info = info.content_size(Some(input.len() as u64));
    }
    let mut enc = FrameEncoder::with_frame_info(info, Vec::new());
    enc.write_all(input).unwrap();
    enc.finish().unwrap()
}

fn xorshift64star(nbytes: usize) -> Vec<u8> {
    let mut x: u64 = 0x9E37_79B9_7F4A_7C15;
    let mut out = Vec::with_capacity(nbytes + 8);
    while out.len() < nbytes {
        x ^= x >> 12;
        x ^= x << 25;
        x ^= x >> 27;
        out.extend_from_slice(&x.wrapping_mul(0x2545_F491_4F6C_DD1D).to_le_bytes());
    }
    out.truncate(nbytes);
    out
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let benches = Path::new(&args[1]);
    let out = Path::new(&args[2]);
    fs::create_dir_all(out).unwrap();
    let load = |f: &str| fs::read(benches.join(f)).unwrap();
    let json = load("compression_66k_JSON.txt");
    let tiled: Vec<u8> = (0..131072).map(|i| json[i % json.len()]).collect();
    let mut inputs: Vec<(String, Vec<u8>)> = Vec::new();
    for f in ["compression_1k.txt", "compression_34k.txt", "compression_65k.txt", "compression_66k_JSON.txt", "dickens.txt"] {
        inputs.push((f.to_string(), load(f)));
    }
    inputs.push(("json_tiled_block0".into(), tiled[..65536].to_vec()));
    inputs.push(("json_tiled_block1".into(), tiled[65536..].to_vec()));
    inputs.push(("zeros_65536".into(), vec![0u8; 65536]));
    inputs.push(("hdfs_first_4MiB".into(), load("hdfs.json")[..4 << 20].to_vec()));
    inputs.push(("xorshift_65536".into(), xorshift64star(65536)));
    for (name, data) in &inputs {
        fs::write(out.join(format!("{name}.block_api.lz4")), lz4_flex::block::compress(data)).unwrap();
        if let Some(b) = frame_block(data, false) {
            fs::write(out.join(format!("{name}.frame_fresh.lz4")), b).unwrap();
        }
        if let Some(b) = frame_block(data, true) {
            fs::write(out.join(format!("{name}.frame_cont.lz4")), b).unwrap();
        }
    }
    let d1m = load("dickens.txt")[..1 << 20].to_vec();
    for (name, data, bsid, flags) in [
        ("json66k_auto", &json, 0u32, 0u32),
        ("json66k_64k", &json, 4, 0),
        ("json66k_64k_all_flags", &json, 4, 7),
        ("dickens1M_64k", &d1m, 4, 0),
        ("dickens1M_256k_checksums", &d1m, 5, 3),
        ("empty_auto", &Vec::new(), 0, 0),
    ] {
        fs::write(out.join(format!("{name}.frame.lz4")), whole_frame(data, bsid, flags)).unwrap();
    }
    println!("wrote crate outputs to {}", out.display());
}
