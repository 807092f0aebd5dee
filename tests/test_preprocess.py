"""Clip preprocessing: uint8 BGR frames -> AlignedResize_clips -> Normalize_clips -> CHW fp32.

CPU: the oracle restatement is pinned bit for bit against OpenCV itself (cv2 is the third-party implementation the
reference's mmcv wrappers call; it is installed in this image).  GPU: the CUDA kernels against the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as PO

cv2 = pytest.importorskip("cv2")
MEAN, STD = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
SIZES = [(480, 853, 480, 864), (480, 853, 480, 853), (37, 53, 64, 96), (360, 640, 480, 864), (100, 100, 73, 131),
         (720, 1280, 480, 864), (200, 300, 100, 150), (50, 70, 96, 160)]


def rand_img(h, w, seed):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8)


def imnormalize_cv2(img, mean, std, to_rgb=True):
    """mmcv.image.photometric.imnormalize_ (mmcv 1.3.0), the same cv2 calls in the same order."""
    img = img.copy().astype(np.float32)
    mean = np.float64(np.asarray(mean, np.float32).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std, np.float32).reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


@pytest.mark.parametrize("h,w,H,W", SIZES)
def test_oracle_resize_equals_cv2(h, w, H, W):
    img = rand_img(h, w, h + w)
    assert np.array_equal(PO.resize_u8(img, W, H), cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR))


def test_oracle_normalize_equals_mmcv_calls():
    img = rand_img(97, 131, 5)
    assert np.array_equal(PO.normalize(img, MEAN, STD, True), imnormalize_cv2(img, MEAN, STD, True))
    assert np.array_equal(PO.normalize(img, MEAN, STD, False), imnormalize_cv2(img, MEAN, STD, False))


def test_oracle_sizes_follow_the_reference_pipeline():
    assert PO.rescale_size((853, 480), (853, 480)) == (853, 480)
    assert PO.rescale_size((1280, 720), (853, 480)) == (853, 480)
    assert PO.rescale_size((640, 360), (853, 480)) == (853, 480)
    assert PO.aligned_size(853, 480) == (864, 480)
    out, shape, sf = PO.preprocess_frame(rand_img(480, 853, 1))
    assert out.shape == (3, 480, 864) and out.dtype == np.float32 and shape == (480, 864, 3)
    assert np.allclose(sf, [864 / 853, 1.0, 864 / 853, 1.0])


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def pre():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vss_cffm_b200 import preprocess
    return preprocess


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,H,W", SIZES)
def test_gpu_resize_u8_bit_exact(pre, h, w, H, W):
    imgs = np.stack([rand_img(h, w, 10 + i) for i in range(3)])
    got = pre.resize_u8(torch.from_numpy(imgs).cuda(), H, W).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], PO.resize_u8(imgs[i], W, H)), f"image {i}"


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,H,W", [(480, 853, 480, 864), (480, 480, 480, 480), (37, 53, 64, 96), (720, 1280, 480, 864)])
@pytest.mark.parametrize("to_rgb", [True, False])
def test_gpu_resize_normalize_bit_exact(pre, h, w, H, W, to_rgb):
    imgs = np.stack([rand_img(h, w, 20 + i) for i in range(2)])
    got = pre.resize_normalize(torch.from_numpy(imgs).cuda(), H, W, MEAN, STD, to_rgb).cpu().numpy()
    assert got.shape == (2, 3, H, W) and got.dtype == np.float32
    for i in range(2):
        ref = PO.normalize(PO.resize_u8(imgs[i], W, H), MEAN, STD, to_rgb).transpose(2, 0, 1)
        assert np.array_equal(got[i], ref), f"image {i}"


@pytest.mark.gpu
def test_gpu_clip_preprocessor_matches_the_pipeline(pre):
    """ClipPreprocessor == AlignedResize_clips + Normalize_clips + ImageToTensor_clips per frame, stacked frame-major."""
    T, B = 4, 2
    clips = [[rand_img(360, 640, 100 + 10 * b + t) for t in range(T)] for b in range(B)]
    pp = pre.ClipPreprocessor(img_scale=(853, 480), size_divisor=32, mean=MEAN, std=STD, to_rgb=True)
    frames, metas = pp(clips)                                                   # (T, B, 3, H, W) on the GPU
    assert tuple(frames.shape) == (T, B, 3, 480, 864) and frames.dtype == torch.float32 and frames.is_cuda
    for b in range(B):
        for t in range(T):
            ref, shape, sf = PO.preprocess_frame(clips[b][t])
            assert np.array_equal(frames[t, b].cpu().numpy(), ref), (b, t)
        assert metas[b]["img_shape"] == (480, 864, 3) and metas[b]["ori_shape"] == (360, 640, 3)
        assert np.array_equal(metas[b]["scale_factor"], sf) and metas[b]["flip"] is False
    with pytest.raises(Exception):
        pre.resize_u8(torch.from_numpy(clips[0][0][None]), 480, 864)            # CPU tensor: no fallback


@pytest.mark.gpu
def test_gpu_pipeline_from_uint8_frames_equals_direct_call(pre):
    """ClipPipeline fed decoded uint8 frames (preprocessing captured in the CUDA graph) == preprocess, then predict_labels."""
    import vss_cffm_b200 as V
    from vss_cffm_b200 import synth
    from vss_cffm_b200.graph import ClipPipeline
    torch.set_grad_enabled(False)
    m = V.build_segmentor(V.model_cfg("b0"))
    synth.fill_module(m, 5)
    m = m.cuda().eval()
    T, B, h, w = 4, 1, 90, 160
    pp = pre.ClipPreprocessor(img_scale=(160, 90), size_divisor=32)
    (_, _), (H, W) = pp.output_size(h, w)
    assert (H, W) == (96, 160)
    g = torch.Generator().manual_seed(9)
    batches = [torch.randint(0, 256, (T, B, h, w, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    metas = pp.metas(B, h, w)
    pipe = ClipPipeline(m, B, T, H, W, metas, rescale=False, preprocessor=pp, src_hw=(h, w))
    outs = [torch.empty(B, H, W, dtype=torch.int64).pin_memory() for _ in range(3)]
    for x, o in zip(batches, outs):
        pipe.submit(x, o)
    pipe.drain()
    for x, o in zip(batches, outs):
        frames = pp.run(x.cuda().view(T * B, h, w, 3), T, B)
        ref = m.labels_from_frames(frames, metas, rescale=False)
        assert torch.equal(ref.cpu(), o)
