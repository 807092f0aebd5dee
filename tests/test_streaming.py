"""Streaming inference with temporal re-use == the stateless segmentor on the reference's clips."""
import pytest
import torch

from vss_cffm_b200 import synth
from vss_cffm_b200.streaming import clip_indices


def test_clip_indices_follow_the_reference_dataset_rule():
    """mmseg/datasets/custom.py:2365-2386."""
    assert [clip_indices(i) for i in range(3)] == [[0], [1], [2]]
    assert clip_indices(3) == [0, 1, 2, 3] and clip_indices(4) == [0, 2, 3, 4] and clip_indices(5) == [0, 2, 4, 5]
    assert clip_indices(6) == [0, 2, 4, 6] and clip_indices(7) == [0, 3, 5, 7] and clip_indices(8) == [0, 3, 6, 8]
    assert clip_indices(9) == [0, 3, 6, 9] and clip_indices(20) == [11, 14, 17, 20]
    assert clip_indices(5, dilation=(-2, -1)) == [3, 4, 5] and clip_indices(1, dilation=(-2, -1)) == [0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams,graph", [(1, False), (2, False), (2, True)])
def test_stream_equals_stateless_on_the_reference_clips(n_streams, graph):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vss_cffm_b200 as V
    from vss_cffm_b200.streaming import VideoStream
    torch.set_grad_enabled(False)
    m = V.build_segmentor(V.model_cfg("b0"))
    synth.fill_module(m, 17)
    m = m.cuda().eval()
    H, W, n_frames = 64, 96, 24 if graph else 13
    video = [synth.synth_array((n_streams, 3, H, W), 500 + i).cuda() for i in range(n_frames)]
    metas = synth.img_metas(n_streams, H, W)
    stream = VideoStream(m, n_streams, graph=graph)
    for i in range(n_frames):
        got = stream.push(video[i])
        ref = m.predict_labels([video[j] for j in clip_indices(i)], metas)
        assert torch.equal(got, ref), f"frame {i} (clip {clip_indices(i)})"
    assert max(stream.kv) == n_frames - 1 and min(stream.kv) >= n_frames - 10      # bounded history
    stream.reset()
    assert torch.equal(stream.push(video[0]), m.predict_labels([video[0]], metas))
