"""CPU tests of the frame-pipeline host logic (depth_pro/pipeline.py): frame selection, the
processing_progress.json resume contract, threaded ingest, sharding, per-frame error isolation.
The GPU side is replaced by a stand-in with the same `infer` signature; the real model runs the same
loop in tests/test_gpu_video.py."""

import json
import os
import threading
import time

import numpy as np
import pytest
import torch

from depth_pro import pipeline


def _touch(d, names):
    for n in names:
        (d / n).write_bytes(b"x")


def test_list_frames_pattern_and_numeric_range(tmp_path):
    _touch(tmp_path, ["output_0001.png", "output_0002.png", "output_0010.png", "output_final.png", "other_0003.png"])
    all_ = pipeline.list_frames(str(tmp_path))
    assert [os.path.basename(p) for p in all_] == ["output_0001.png", "output_0002.png", "output_0010.png",
                                                    "output_final.png"]
    # a range drops names without digits and filters inclusively on the number in the name
    rng = pipeline.list_frames(str(tmp_path), start_frame=2, end_frame=10)
    assert [os.path.basename(p) for p in rng] == ["output_0002.png", "output_0010.png"]
    assert [os.path.basename(p) for p in pipeline.list_frames(str(tmp_path), end_frame=1)] == ["output_0001.png"]
    assert pipeline.frame_number("a/b/clip2_frame_0243.png") == 20243  # all digits of the base name, like the reference
    assert pipeline.frame_number("nodigits.png") is None


def test_progress_file_format_resume_and_force(tmp_path):
    out = str(tmp_path)
    paths = [os.path.join(out, f"output_{i:04d}.png") for i in range(6)]
    p = pipeline.Progress(out, save_every=2)
    p.mark(paths[0], True)
    assert not os.path.exists(p.path)           # saved every 2 successes
    p.mark(paths[1], False)                     # failures are not recorded
    p.mark(paths[2], True)
    data = json.load(open(p.path))
    assert set(data) == {"output_0000.png", "output_0002.png"}
    assert data["output_0000.png"]["success"] is True and isinstance(data["output_0000.png"]["timestamp"], float)
    # resume skips completed frames, keeps order; without resume / with force nothing is skipped
    r = pipeline.Progress(out, resume=True)
    assert [os.path.basename(x) for x in r.pending(paths)] == ["output_0001.png", "output_0003.png", "output_0004.png",
                                                                "output_0005.png"]
    assert len(pipeline.Progress(out, resume=False).pending(paths)) == 6
    assert len(pipeline.Progress(out, resume=True, force_reprocess=True).pending(paths)) == 6
    # a corrupt progress file counts as empty
    open(p.path, "w").write("{not json")
    assert len(pipeline.Progress(out, resume=True).pending(paths)) == 6


def test_progress_shards_merge(tmp_path):
    out = str(tmp_path)
    a = pipeline.Progress(out, rank=0, world=2, save_every=1)
    b = pipeline.Progress(out, rank=1, world=2, save_every=1)
    a.mark("f0.png", True), b.mark("f1.png", True), a.mark("f2.png", True)
    assert os.path.exists(os.path.join(out, "processing_progress.rank0.json"))
    assert os.path.exists(os.path.join(out, "processing_progress.rank1.json"))
    merged = a.merge_shards()
    assert set(merged) == {"f0.png", "f1.png", "f2.png"}
    assert set(json.load(open(os.path.join(out, pipeline.PROGRESS_FILE)))) == set(merged)
    # a later resumed run on a different world size sees the merged file
    assert pipeline.Progress(out, resume=True, rank=2, world=4).pending(["x/f1.png", "x/f9.png"]) == ["x/f9.png"]


def test_prepare_image_downscale_semantics():
    img = (np.arange(40 * 60 * 3) % 251).astype(np.uint8).reshape(40, 60, 3)
    same, f = pipeline.prepare_image(img, 100.0, 1.0)
    assert same.shape == (40, 60, 3) and f == 100.0
    half, f = pipeline.prepare_image(img, 100.0, 0.5)
    assert half.shape == (20, 30, 3) and f == 50.0 and half.flags["C_CONTIGUOUS"]
    odd, f = pipeline.prepare_image(img, None, 0.33)
    assert odd.shape == (int(40 * 0.33), int(60 * 0.33), 3) and f is None
    up, f = pipeline.prepare_image(img, 10.0, 2.0)
    assert up.shape == (80, 120, 3) and f == 20.0
    import cv2

    assert np.array_equal(half, cv2.resize(img, (30, 20), interpolation=cv2.INTER_AREA))


def test_frame_loader_order_prefetch_and_errors():
    active, peak, lock = [0], [0], threading.Lock()

    def load(path):
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        time.sleep(0.01 if path.endswith("3") else 0.002)
        with lock:
            active[0] -= 1
        if path.endswith("5"):
            raise OSError("cannot identify image file")
        i = int(path[1:])
        return np.full((4, 6, 3), i, np.uint8), None, 10.0 * i

    items = [(i, f"f{i}") for i in range(12)]
    frames = list(pipeline.FrameLoader(items, downscale_factor=0.5, threads=4, prefetch=6, load_fn=load))
    assert [f.index for f in frames] == list(range(12))                     # clip order, whatever finishes first
    assert frames[5].image is None and "cannot identify" in frames[5].error
    assert frames[7].image.shape == (2, 3, 3) and int(frames[7].image[0, 0, 0]) == 7 and frames[7].f_px == 35.0
    assert 1 < peak[0] <= 4                                                 # decoding really ran in parallel


class _FakeModel:
    """Same `infer` contract as DepthPro.infer for uint8 HWC input; depth = mean intensity + 1."""

    def __init__(self):
        self.calls = 0

    def infer(self, x, f_px=None):
        self.calls += 1
        x = torch.from_numpy(x) if isinstance(x, np.ndarray) else x
        h, w, _ = x.shape
        return {"depth": torch.full((h, w), float(x.float().mean()) + 1.0),
                "focallength_px": torch.tensor(float(f_px) if f_px is not None else 0.5 * w)}


def _make_clip(d, n):
    from PIL import Image

    for i in range(n):
        Image.fromarray(np.full((8, 12, 3), 10 * i, np.uint8)).save(d / f"output_{i:04d}.png")
    (d / "output_9999.png").write_bytes(b"not a png")


def test_process_frames_single_inference_resume_and_isolation(tmp_path):
    frames, out = tmp_path / "frames", tmp_path / "out"
    frames.mkdir()
    _make_clip(frames, 5)
    model, seen, logs = _FakeModel(), [], []

    def consumer(o):
        seen.append((o.index, os.path.basename(o.path), float(o.depth[0, 0]), o.focallength_px, o.image.shape))
        return o.index != 3                                  # frame 3 reports failure

    s = pipeline.process_frames(str(frames), str(out), model, consumer, unproject=False, decode_threads=2, log=logs.append)
    assert s.total == 6 and s.skipped == 0 and s.processed == 4
    assert sorted(s.failed) == ["output_0003.png", "output_9999.png"]
    assert model.calls == 5                                  # ONE inference per decodable frame, none for the broken file
    assert [x[0] for x in seen] == [0, 1, 2, 3, 4] and seen[2][2] == 21.0 and seen[2][3] == 6.0
    assert any("output_9999.png" in m for m in logs)
    done = json.load(open(out / pipeline.PROGRESS_FILE))
    assert set(done) == {"output_0000.png", "output_0001.png", "output_0002.png", "output_0004.png"}

    # resume: only the two unfinished frames are retried
    seen.clear()
    s2 = pipeline.process_frames(str(frames), str(out), model, lambda o: seen.append(o.index), unproject=False,
                                 resume=True, log=logs.append)
    assert s2.skipped == 4 and s2.processed == 1 and s2.failed == ["output_9999.png"] and model.calls == 6
    # downscale reaches the network input and the focal length
    seen.clear()
    pipeline.process_frames(str(frames), str(out), model, lambda o: seen.append(o.image.shape), unproject=False,
                            downscale_factor=0.5, end_frame=1, log=logs.append)
    assert seen == [(4, 6, 3), (4, 6, 3)]
    assert pipeline.process_frames(str(frames), str(out), model, consumer, pattern="nothing_*.png", log=logs.append).total == 0


def test_process_frames_shards_partition_the_clip(tmp_path):
    frames = tmp_path / "frames"
    frames.mkdir()
    _make_clip(frames, 7)
    got = {}
    for rank in (1, 0):      # rank 0 last: it merges the shard files (no process group in this test)
        seen = []
        pipeline.process_frames(str(frames), str(tmp_path / "out"), _FakeModel(), lambda o: seen.append(o.index),
                                unproject=False, pattern="output_000*.png", rank=rank, world=2, log=lambda m: None)
        got[rank] = seen
    assert got[0] == [0, 2, 4, 6] and got[1] == [1, 3, 5]
    assert len(json.load(open(tmp_path / "out" / pipeline.PROGRESS_FILE))) == 7


def test_interrupted_multi_rank_resume_partitions_exactly_the_pending_set(tmp_path):
    """after a killed world=2 run there are two rank shard files and NO merged file.  On resume
    every rank must see the union of what both ranks finished, and the frames still to do must be split by CLIP
    POSITION (i % world), so that the two ranks together process exactly the pending set, nothing twice."""
    frames, out = tmp_path / "frames", tmp_path / "out"
    frames.mkdir(), out.mkdir()
    _make_clip(frames, 10)
    stamp = {"success": True, "timestamp": 1.0}
    # the interrupted run: rank 0 had finished clip positions 0, 2, 4; rank 1 had finished 1, 3
    json.dump({f"output_{i:04d}.png": stamp for i in (0, 2, 4)}, open(out / "processing_progress.rank0.json", "w"))
    json.dump({f"output_{i:04d}.png": stamp for i in (1, 3)}, open(out / "processing_progress.rank1.json", "w"))
    got = {}
    for rank in (1, 0):
        seen = []
        s = pipeline.process_frames(str(frames), str(out), _FakeModel(), lambda o: seen.append(o.index), unproject=False,
                                    pattern="output_000*.png", resume=True, rank=rank, world=2, log=lambda m: None)
        got[rank] = seen
        # the ranks run one after the other here: rank 0 also sees the three frames rank 1 has just finished
        assert s.total == 10 and s.skipped == (5 if rank == 1 else 8)
    assert got[0] == [6, 8] and got[1] == [5, 7, 9]                         # clip positions, owner = i % 2
    assert sorted(got[0] + got[1]) == [5, 6, 7, 8, 9]                       # exactly the pending set, no overlap
    merged = json.load(open(out / pipeline.PROGRESS_FILE))
    assert set(merged) == {f"output_{i:04d}.png" for i in range(10)}


def test_pin_to_gpu_numa_is_safe_without_a_gpu():
    """No NVML / no GPU: placement must be a silent no-op that leaves the affinity mask alone."""
    before = os.sched_getaffinity(0)
    got = pipeline.pin_to_gpu_numa(0)
    assert got is None or set(got) <= set(before)
    if got is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
