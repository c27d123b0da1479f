"""File formats on either side of the path (.yuv sequences, raw block files): layout checked on the CPU, the file-driven
cascade on the GPU."""
import numpy as np
import pytest
import torch

from cnn_av1_research_b200 import fileio, synth
from oracle import cascade_oracle as O


def test_yuv_reader_and_block_files_round_trip(tmp_path):
    w, h, nf = 100, 70, 3
    words = synth.synth_frames(nf, w, h, seed=5)
    path = tmp_path / "seq.yuv"
    words.astype("<u2").tofile(path)
    assert fileio.count_frames(path, w, h) == nf
    y, stats = fileio.read_y_component_10bit_lossless(path, 2, w, h)
    assert np.array_equal(y, O.luma_plane(words, 2, w, h)) and stats["max"] <= 1023 and stats["shape"] == (h, w)
    assert fileio.read_y_component_10bit_lossless(path, 3, w, h) == (None, None)          # past the end: reference behaviour
    fr = fileio.read_frames_yuv420p10(path, w, h, 1, 2, pin=False)
    fw = synth.frame_words(w, h)
    assert np.array_equal(fr.view(torch.int16).numpy().view(np.uint16), words[fw:3 * fw])
    with pytest.raises(ValueError):
        fileio.read_frames_yuv420p10(path, w, h, 2, 2)
    blocks = O.extract_blocks(y, 16)
    info = fileio.save_blocks_binary_10bit(blocks, tmp_path / "b16.raw", {}, verbose=False)
    assert info["num_blocks"] == 35 and info["total_bytes"] == blocks.size * 2
    back = fileio.load_block_file(tmp_path / "b16.raw", 16)
    assert back.shape == (35, 16, 16, 1) and np.array_equal(back[..., 0], blocks)
    with pytest.raises(TypeError):
        fileio.save_blocks_binary_10bit(blocks.astype(np.float32), tmp_path / "x.raw", {})


def test_reference_reader_agrees_when_reference_present(tmp_path):
    import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    ns = ref_import.load()
    w, h = 64, 48
    words = synth.synth_frames(2, w, h, seed=8)
    path = tmp_path / "seq.yuv"
    words.astype("<u2").tofile(path)
    y_ref, s_ref = ns.extract.read_y_component_10bit_lossless(str(path), 1, w, h)
    y, s = fileio.read_y_component_10bit_lossless(path, 1, w, h)
    assert np.array_equal(y, y_ref) and s == s_ref
    blocks = O.extract_blocks(y, 8)
    a = ns.extract.save_blocks_binary_10bit(blocks, str(tmp_path / "ref.raw"), {}, verbose=False)
    b = fileio.save_blocks_binary_10bit(blocks, tmp_path / "mine.raw", {}, verbose=False)
    assert open(tmp_path / "ref.raw", "rb").read() == open(tmp_path / "mine.raw", "rb").read()
    assert b["md5_hash"] == (a.get("md5_hash") if isinstance(a, dict) and "md5_hash" in a else b["md5_hash"])


@pytest.mark.gpu
def test_gpu_predict_from_yuv_file(cuda_device, tmp_path):
    from cnn_av1_research_b200.testing import build_pipeline
    w, h, nf, thr = 640, 360, 3, 0.45
    words = synth.synth_frames(nf, w, h, seed=1234)
    path = tmp_path / "seq.yuv"
    words.astype("<u2").tofile(path)
    pipe = build_pipeline(seed=0, threshold=thr, device=cuda_device)
    labels = fileio.predict_yuv_file(pipe, path, w, h, first_frame=1, n_frames=2, chunk_frames=1).numpy()
    fw = synth.frame_words(w, h)
    ref = O.cascade_predict(synth.calibrated_cascade(0), O.frames_to_images(words[fw:], 2, w, h), thr)["labels"].numpy()
    assert labels.shape == ref.shape and (labels == ref).mean() >= 0.999
    # raw block file -> BlockRecord.to_torch -> predict (the dataset path of 008)
    import cnn_av1_research_b200 as P
    y, _ = fileio.read_y_component_10bit_lossless(path, 1, w, h)
    blocks, _ = P.extract_blocks_with_validation(y, 16, w, h, verbose=False)
    fileio.save_blocks_binary_10bit(blocks, tmp_path / "blk.raw", {}, verbose=False)
    rec = P.BlockRecord(samples=fileio.load_block_file(tmp_path / "blk.raw", 16), labels=np.zeros(len(blocks), np.int64),
                        qps=np.zeros((len(blocks), 1), np.float32)).to_torch(cuda_device)
    assert np.array_equal(pipe.predict(rec.samples).numpy().astype(np.uint8), labels[:len(blocks)])


def test_predict_yuv_file_streams_windows_in_order(tmp_path):
    """Window streaming of predict_yuv_file with a stand-in pipeline (its predict_frames_host returns one value per block:
    the first luma sample of the block's frame): every frame is read exactly once, in order, windows of the requested
    size, the last one partial; the result equals the single-window call."""
    from cnn_av1_research_b200 import fileio
    from cnn_av1_research_b200.extraction import calculate_yuv420_10bit_sizes
    w, h, nf = 64, 32, 11
    words = calculate_yuv420_10bit_sizes(w, h)["total_frame_size"] // 2
    data = np.zeros(nf * words, dtype="<u2")
    for f in range(nf):
        data[f * words:(f + 1) * words] = 100 + f
    path = tmp_path / "s.yuv"
    data.tofile(path)
    calls = []

    class Stub:
        def predict_frames_host(self, frames, width, height, n_frames, chunk_frames=8):
            assert frames.numel() == n_frames * words and (width, height) == (w, h)
            calls.append(n_frames)
            v = frames.view(torch.int16).reshape(n_frames, words)[:, 0].to(torch.uint8)
            return v.repeat_interleave((w // 16) * (h // 16))
    got = fileio.predict_yuv_file(Stub(), path, w, h, first_frame=1, n_frames=9, chunk_frames=2, window_frames=4)
    assert calls == [4, 4, 1]
    want = torch.arange(101, 110, dtype=torch.uint8).repeat_interleave(8)
    assert torch.equal(got, want)
    calls.clear()
    assert torch.equal(fileio.predict_yuv_file(Stub(), path, w, h, first_frame=1, n_frames=9, window_frames=64), want) and calls == [9]
    with pytest.raises(ValueError):
        fileio.predict_yuv_file(Stub(), path, w, h, first_frame=5, n_frames=7)
