"""The assembled north-star path -- ``DualHUSynthesizer`` (reference generate.py:89-102 + generate.py:213-237) -- at BASELINE
size on the B200 against the oracle: 512x512 slices, both generators with 9 CBAM blocks, a 33-slice volume run as 17 + 16
slices (batch_slices = 20: a full chunk + a shorter tail through the two-stream, chunk-pipelined path that bench.py times),
as 7+7+7+7+5, as 3 x 11 and as one batch of 33 (synthesis.chunk_size).

Gates: voxels outside both HU ranges keep the raw stored value bit for bit; slice i of the output is slice i of the input;
generator-derived voxels within the stated bound (tanh tolerance 0.015 of tests/test_gpu_generator.py = 3 HU in the 400 HU
soft-tissue window, 6.4 HU in the 850 HU lung window, +1 for the truncating cast of preprocess.py:111); the device-resident
entry, other chunkings and repeated calls agree bit for bit; postprocess=True equals the scipy pipeline on the same merged
volume bit for bit.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402

S, H, W, NB = 33, 512, 512, 9
SLOPE, INTERCEPT = 1.0, -1024.0
MAX_SOFT, MAX_LUNG = 4, 8            # stored units (= HU at slope 1): 0.015 tanh units of each window + 1 (truncation)
MEAN_SOFT, MEAN_LUNG = 0.6, 1.2      # 0.002 tanh units (TOL_MEAN of test_gpu_generator.py) of each window + truncation bias


@pytest.fixture(scope="module")
def case():
    shapes = orc.generator_param_shapes(1, NB, True)
    sd_s, sd_l = orc.make_state_dict(shapes, 1234), orc.make_state_dict(shapes, 1235)
    # 20 slices of the bench's uniform 0..2500 volume (every mask branch in every neighbourhood) + 13 phantom slices
    vol = np.concatenate([orc.synthetic_volume(20, H, W, seed=3), orc.phantom_volume(13, H, W, seed=1)]).astype(np.int16)
    assert vol.shape == (S, H, W)
    ref = orc.dual_hu_synthesize(vol, SLOPE, INTERCEPT, sd_s, sd_l, NB, True)
    return sd_s, sd_l, vol, ref


def _synth(sd_s, sd_l, batch_slices):
    from ducosy_gan_b200.modules.model import Generator
    from ducosy_gan_b200.synthesis import DualHUSynthesizer
    gs, gl = Generator(1, NB), Generator(1, NB)
    gs.load_state_dict(sd_s, strict=True)
    gl.load_state_dict(sd_l, strict=True)
    return DualHUSynthesizer(gs.cuda().eval(), gl.cuda().eval(), batch_slices=batch_slices)


def _check_against_oracle(merged, vol, ref, what):
    hu = orc.stored_to_hu(vol, SLOPE, INTERCEPT)
    soft = (hu >= -150) & (hu <= 250)
    lung = (hu >= -1000) & (hu <= -150)            # lung wins at HU == -150 (generate.py:229-232)
    soft_only = soft & ~lung
    outside = ~(soft | lung)
    assert merged.dtype == np.int16 and merged.shape == vol.shape
    assert np.array_equal(merged[outside], vol[outside]), f"{what}: raw voxels outside both HU ranges must be kept bit for bit"
    d = np.abs(merged.astype(np.int32) - ref.astype(np.int32))
    per_slice = d.reshape(S, -1).max(axis=1)
    print(f"{what}: soft-window voxels max {d[soft_only].max()} mean {d[soft_only].mean():.3f}; lung-window voxels max "
          f"{d[lung].max()} mean {d[lung].mean():.3f} (stored units = HU); worst slice {int(per_slice.argmax())}")
    # a slice written to the wrong index (or a chunk hand-off race) shows up as hundreds of HU, not as rounding noise
    assert d[soft_only].max() <= MAX_SOFT and d[lung].max() <= MAX_LUNG, (d[soft_only].max(), d[lung].max())
    assert d[soft_only].mean() <= MEAN_SOFT and d[lung].mean() <= MEAN_LUNG


def test_synthesize_volume_33_slices_two_chunks_vs_oracle(case):
    """the call bench.py's e2e leg times: pinned host volume in, pinned host volume out, a 17-slice chunk + a 16-slice tail"""
    from ducosy_gan_b200.synthesis import chunk_size
    sd_s, sd_l, vol, ref = case
    assert chunk_size(S, 20) == 17 and chunk_size(S, 30) == 33 and chunk_size(S, 7) == 7 and chunk_size(300, 30) == 30
    synth = _synth(sd_s, sd_l, 20)
    host = torch.from_numpy(vol).pin_memory()
    merged = synth.synthesize_volume(host, SLOPE, INTERCEPT)
    assert merged.device.type == "cpu"
    _check_against_oracle(merged.numpy(), vol, ref, "synthesize_volume")
    again = synth.synthesize_volume(host, SLOPE, INTERCEPT)          # reused ys / yl / workspaces / streams
    assert torch.equal(merged, again)
    assert np.array_equal(synth.synthesize_volume(vol, SLOPE, INTERCEPT).numpy(), merged.numpy())      # pageable numpy input


def test_synthesize_device_and_other_chunkings_agree_bit_for_bit(case):
    """the device-resident entry (bench.py's `value` leg) and different chunkings (17+16, 3 x 11, 33 at once, 7-slice chunks
    with a 5-slice tail): the kernels are deterministic and batch invariant, so every variant must give the same bytes"""
    sd_s, sd_l, vol, ref = case
    dev = torch.from_numpy(vol).cuda()
    base = _synth(sd_s, sd_l, 20).synthesize_device(dev, SLOPE, INTERCEPT)
    _check_against_oracle(base.cpu().numpy(), vol, ref, "synthesize_device")
    for bs in (16, 30, 7):
        synth = _synth(sd_s, sd_l, bs)
        out = torch.empty_like(dev)
        got = synth.synthesize_device(dev, SLOPE, INTERCEPT, out=out)
        assert got is out and torch.equal(out, base), bs
    # a shard of the volume (what rank r of a slice-sharded run computes) is the same bytes as those slices of the whole
    from ducosy_gan_b200.synthesis import shard_range
    synth = _synth(sd_s, sd_l, 30)
    for r in range(4):
        lo, hi = shard_range(S, r, 4)
        assert torch.equal(synth.synthesize_device(dev[lo:hi].contiguous(), SLOPE, INTERCEPT), base[lo:hi]), r


def test_synthesize_with_postprocess_equals_scipy_on_the_merged_volume(case):
    """generate.py:254-263 appended on the device: bit-exact against the oracle's scipy pipeline applied to the same
    merged volume (the smoothing is exact integer/float64 work; the generator tolerance lives in the merged volume)"""
    sd_s, sd_l, vol, _ = case
    synth = _synth(sd_s, sd_l, 30)
    host = torch.from_numpy(vol).pin_memory()
    merged = synth.synthesize_volume(host, SLOPE, INTERCEPT).numpy().copy()
    smoothed = synth.synthesize_volume(host, SLOPE, INTERCEPT, postprocess=True).numpy()
    assert np.array_equal(smoothed, orc.postprocess_volume(merged))
    dev = synth.synthesize_device(torch.from_numpy(vol).cuda(), SLOPE, INTERCEPT, postprocess=True)
    assert np.array_equal(dev.cpu().numpy(), smoothed)


def test_synthesize_edge_sizes(case):
    sd_s, sd_l, vol, ref = case
    synth = _synth(sd_s, sd_l, 30)
    empty = synth.synthesize_device(torch.empty((0, H, W), dtype=torch.int16, device="cuda"))
    assert empty.shape == (0, H, W)
    assert synth.synthesize_volume(np.empty((0, H, W), np.int16)).shape == (0, H, W)
    one = synth.synthesize_volume(vol[5:6], SLOPE, INTERCEPT).numpy()
    d = np.abs(one.astype(np.int32) - ref[5:6].astype(np.int32))
    assert d.max() <= MAX_LUNG
    with pytest.raises(RuntimeError):
        synth.synthesize_volume(vol.astype(np.float32))


def test_synthesize_series_dicom_folder_round_trip(tmp_path):
    """NCCT series folder -> synthetic series folder through ``dicom_io.synthesize_series`` (one read and one
    write per slice instead of generate.py's per-slice DICOM round trips) equals ``synthesize_volume`` (+ the device volume
    smoothing) on the same stored values; files are named {idx:04d}.dcm in sorted-filename order (generate.py:88,285)."""
    from test_dicom_io import _file      # tests/ is on sys.path (pytest rootdir import mode)
    from ducosy_gan_b200 import dicom_io as dio
    from ducosy_gan_b200.modules.model import Generator
    from ducosy_gan_b200.postprocess import postprocess_volume
    from ducosy_gan_b200.synthesis import DualHUSynthesizer
    nb = 1
    shapes = orc.generator_param_shapes(1, nb, True)
    gs, gl = Generator(1, nb), Generator(1, nb)
    gs.load_state_dict(orc.make_state_dict(shapes, 3))
    gl.load_state_dict(orc.make_state_dict(shapes, 4))
    synth = DualHUSynthesizer(gs.cuda().eval(), gl.cuda().eval(), batch_slices=4)
    vol = orc.phantom_volume(7, H, W, seed=2).astype(np.int16)
    src = tmp_path / "POST VUE"
    src.mkdir()
    for i in range(7):
        (src / f"IM{i:03d}.dcm").write_bytes(_file(vol[i], explicit=i % 2 == 0, slope="1", intercept="-1024"))
    merged, slices = dio.synthesize_series(synth, str(src), str(tmp_path / "out"), postprocess=True)
    want = synth.synthesize_volume(torch.from_numpy(vol), 1.0, -1024.0)
    want = postprocess_volume(want.cuda()).cpu()
    assert torch.equal(merged, want)
    for i in range(7):
        back = dio.read_dicom(str(tmp_path / "out" / f"{i:04d}.dcm"))
        assert np.array_equal(back.pixel_array(), want[i].numpy())
        assert back.series_description == "DuCoSyGAN sCECT v2" and back.transfer_syntax == dio.EXPLICIT_LE
