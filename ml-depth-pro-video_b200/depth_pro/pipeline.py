"""Frame-pipeline orchestration around the hot path.

The reference drives ``model.infer`` from ``pointcloud_pipeline.process_images_to_floor_plans``
(pointcloud_pipeline.py:473-771): glob + numeric frame-range filter (:524-548), a resumable
``processing_progress.json`` (:561-582, :615-622), a pool of CPU worker processes that all share
``cuda:0`` and rebuild the model per frame (:629-714, generate_depth_maps.py:76-80), two inferences
per frame (ground fit + depth), and serial image decode.  This module keeps that contract -- same
frame selection, same progress-file format, per-frame error isolation -- and replaces the execution
model with what a B200 box wants:

  * one process per GPU, the model built ONCE per process, frame ``i`` -> rank ``i % world``;
  * decode + EXIF focal length + ``downscale_factor`` (cv2 INTER_AREA / INTER_LINEAR,
    generate_depth_maps.py:91-110) on a pool of host threads that runs ahead of the GPU;
  * ONE ``infer`` per frame; the optional depth -> 3-D unprojection (+ colours) runs on the GPU
    right behind it; whatever comes next (ground fit, meshing, floor plan: CPU geometry, out of
    scope) plugs in as ``consumer(result)``;
  * every rank keeps its own ``processing_progress.rank<r>.json`` (no cross-process file races);
    rank 0 merges them into ``processing_progress.json`` at the end of the clip.

Only host logic lives here; every tensor operation goes through ``DepthPro.infer`` /
``video.depth_to_3d`` (C-ABI, CUDA).  There is no CPU inference path.
"""

from __future__ import annotations

import glob
import json
import os
import time
from collections import deque
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

PROGRESS_FILE = "processing_progress.json"


# ----------------------------------------------------------------------------------------------
# host placement: one process per GPU wants its threads and pinned buffers on the GPU's NUMA node
# ----------------------------------------------------------------------------------------------
def pin_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Restrict this process (and the threads / pinned allocations it makes from now on: first touch) to the CPU cores
    NVML reports as local to GPU ``device_index``.  The reference's worker pool leaves placement to the OS with every
    worker on ``cuda:0`` (pointcloud_pipeline.py:629-714); with one process per GPU on a two-socket B200 box a rank
    whose decode / copy threads sit on the far socket pays for every H2D / D2H byte twice.  Returns the core list, or
    None when NVML or ``sched_setaffinity`` is unavailable or the mask is empty (nothing is changed then).  If
    ``CUDA_VISIBLE_DEVICES`` remaps devices, pass the physical index."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 -- placement is an optimisation, never a requirement
        return None


# ----------------------------------------------------------------------------------------------
# frame selection  (pointcloud_pipeline.py:524-548)
# ----------------------------------------------------------------------------------------------
def frame_number(path: str) -> Optional[int]:
    """All digits of the file name read as one integer (``output_0243.png`` -> 243); None if there are none."""
    digits = "".join(ch for ch in os.path.basename(path) if ch.isdigit())
    return int(digits) if digits else None


def list_frames(frames_dir: str, pattern: str = "output_*.png", start_frame: Optional[int] = None,
                end_frame: Optional[int] = None) -> List[str]:
    """Sorted glob of ``frames_dir/pattern``; with a range, files without a number are dropped."""
    paths = sorted(glob.glob(os.path.join(frames_dir, pattern)))
    if start_frame is None and end_frame is None:
        return paths
    keep = []
    for p in paths:
        n = frame_number(p)
        if n is None:
            continue
        if (start_frame is None or n >= start_frame) and (end_frame is None or n <= end_frame):
            keep.append(p)
    return keep


# ----------------------------------------------------------------------------------------------
# resume file  (pointcloud_pipeline.py:561-582, 615-622, 766-768)
# ----------------------------------------------------------------------------------------------
class Progress:
    """``{basename: {"success": bool, "timestamp": float}}`` -- the reference's progress-file format.

    ``resume`` loads what earlier runs completed (an unreadable file counts as empty, like the
    reference); ``force_reprocess`` ignores it.  Writes go to a temporary file first and are renamed
    into place, so a kill mid-write never leaves a truncated JSON behind.
    """

    def __init__(self, output_dir: str, resume: bool = False, force_reprocess: bool = False, rank: int = 0,
                 world: int = 1, save_every: int = 5):
        self.output_dir, self.rank, self.world, self.save_every = output_dir, rank, world, max(1, save_every)
        self.path = os.path.join(output_dir, PROGRESS_FILE)
        self.shard_path = self.path if world == 1 else os.path.join(output_dir, f"processing_progress.rank{rank}.json")
        self.done: Dict[str, Dict[str, Any]] = {}
        self._unsaved = 0
        if resume and not force_reprocess:
            # EVERY rank reads the merged file and ALL rank shards: after an interrupted multi-rank run the merged file
            # may be missing or stale, and a rank that only knew its own shard would disagree with the others about
            # what is left (frames skipped or processed twice).
            self.done.update(self._read(self.path))
            for p in sorted(glob.glob(os.path.join(output_dir, "processing_progress.rank*.json"))):
                self.done.update(self._read(p))

    @staticmethod
    def _read(path: str) -> Dict[str, Dict[str, Any]]:
        try:
            with open(path) as f:
                data = json.load(f)
            return data if isinstance(data, dict) else {}
        except (OSError, ValueError):
            return {}

    def pending(self, paths: Sequence[str]) -> List[str]:
        """Frames still to do, in order (only successful entries are skipped on resume)."""
        return [p for p in paths if not self.done.get(os.path.basename(p), {}).get("success", False)]

    def mark(self, path: str, success: bool) -> None:
        if success:  # the reference records successes only (:694, :744)
            self.done[os.path.basename(path)] = {"success": True, "timestamp": time.time()}
            self._unsaved += 1
            if self._unsaved >= self.save_every:
                self.save()

    def save(self, path: Optional[str] = None) -> None:
        path = path or self.shard_path
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "w") as f:
            json.dump(self.done, f, indent=2)
        os.replace(tmp, path)
        self._unsaved = 0

    def merge_shards(self) -> Dict[str, Dict[str, Any]]:
        """Rank 0, end of clip: fold every ``processing_progress.rank*.json`` into the main file."""
        merged = dict(self._read(self.path))
        merged.update(self.done)
        for p in sorted(glob.glob(os.path.join(self.output_dir, "processing_progress.rank*.json"))):
            merged.update(self._read(p))
        self.done = merged
        self.save(self.path)
        return merged


# ----------------------------------------------------------------------------------------------
# ingest  (utils.py:47-112 load_rgb, generate_depth_maps.py:91-110 downscale)
# ----------------------------------------------------------------------------------------------
@dataclass
class Frame:
    index: int                   # position in the clip (after filtering), the unit that is sharded
    path: str
    image: Optional[np.ndarray]  # uint8 HWC, None if decoding failed
    f_px: Optional[float]
    error: Optional[str] = None


def prepare_image(image: np.ndarray, f_px: Optional[float], downscale_factor: float
                  ) -> Tuple[np.ndarray, Optional[float]]:
    """``downscale_factor`` semantics of generate_depth_maps.py:91-110: size = int(size * factor), INTER_AREA when
    shrinking, INTER_LINEAR when enlarging, the EXIF focal length scales with the image."""
    if downscale_factor != 1.0 and downscale_factor > 0:
        import cv2

        h, w = image.shape[:2]
        nh, nw = int(h * downscale_factor), int(w * downscale_factor)
        image = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_AREA if downscale_factor < 1.0 else cv2.INTER_LINEAR)
        if f_px is not None:
            f_px = f_px * downscale_factor
    return np.ascontiguousarray(image), f_px


class FrameLoader:
    """Decodes frames on ``threads`` host threads, at most ``prefetch`` frames ahead of the consumer, and yields
    them in clip order.  A frame that fails to decode is yielded with ``image=None`` and the message (the
    reference prints the error and moves on to the next frame, generate_depth_maps.py:147-151)."""

    def __init__(self, items: Sequence[Tuple[int, str]], downscale_factor: float = 1.0, threads: int = 4,
                 prefetch: int = 8, load_fn: Optional[Callable[[str], Tuple[np.ndarray, Any, Optional[float]]]] = None):
        if load_fn is None:
            from .utils import load_rgb as load_fn
        self.items, self.factor, self.load_fn = list(items), downscale_factor, load_fn
        self.threads, self.prefetch = max(1, threads), max(1, prefetch)

    def _load(self, index: int, path: str) -> Frame:
        try:
            image, _, f_px = self.load_fn(path)
            image, f_px = prepare_image(image, f_px, self.factor)
            return Frame(index, path, image, f_px)
        except Exception as e:  # noqa: BLE001 -- per-frame isolation
            return Frame(index, path, None, None, f"{type(e).__name__}: {e}")

    def __iter__(self) -> Iterator[Frame]:
        queue: "deque[Future]" = deque()
        with ThreadPoolExecutor(max_workers=self.threads, thread_name_prefix="depthpro-decode") as pool:
            it = iter(self.items)
            for index, path in it:
                queue.append(pool.submit(self._load, index, path))
                if len(queue) >= self.prefetch:
                    break
            while queue:
                frame = queue.popleft().result()
                nxt = next(it, None)
                if nxt is not None:
                    queue.append(pool.submit(self._load, *nxt))
                yield frame


# ----------------------------------------------------------------------------------------------
# the loop
# ----------------------------------------------------------------------------------------------
@dataclass
class FrameOutput:
    index: int
    path: str
    depth: Any                       # (H,W) float32 CUDA tensor, metres
    focallength_px: float
    image: np.ndarray                # uint8 HWC as fed to the network (after downscale)
    points: Any = None               # (N,3) float32 CUDA tensor if unproject=True
    colors: Any = None               # (N,3) float32 CUDA tensor in [0,1] if unproject=True
    valid_mask: Any = None


@dataclass
class ClipSummary:
    total: int = 0                   # frames selected by pattern / range
    skipped: int = 0                 # already completed (resume)
    processed: int = 0               # successful on this rank
    failed: List[str] = field(default_factory=list)
    seconds: float = 0.0


def process_frames(frames_dir: str, output_dir: Optional[str], model, consumer: Callable[[FrameOutput], Any],
                   pattern: str = "output_*.png", start_frame: Optional[int] = None, end_frame: Optional[int] = None,
                   downscale_factor: float = 1.0, unproject: bool = True, resume: bool = False,
                   force_reprocess: bool = False, rank: int = 0, world: int = 1, decode_threads: int = 4,
                   load_fn: Optional[Callable] = None, log: Callable[[str], None] = print) -> ClipSummary:
    """Run the hot path over a directory of frames: select -> (resume filter) -> shard -> threaded decode ->
    ONE ``model.infer`` per frame -> optional GPU unprojection -> ``consumer`` -> progress file.

    ``consumer`` receives a :class:`FrameOutput` whose tensors live on the model's GPU; returning ``False``
    marks the frame as failed (anything else, including ``None``, is success).  Exceptions raised while a
    frame is processed are reported and the loop continues, like the reference's workers (:398-470).
    """
    from . import video

    output_dir = output_dir or frames_dir
    os.makedirs(output_dir, exist_ok=True)
    t0 = time.time()
    paths = list_frames(frames_dir, pattern, start_frame, end_frame)
    summary = ClipSummary(total=len(paths))
    if not paths:
        log(f"No images found matching pattern {os.path.join(frames_dir, pattern)}")
        return summary
    progress = Progress(output_dir, resume, force_reprocess, rank, world)
    todo = progress.pending(paths) if (resume and not force_reprocess) else paths
    summary.skipped = len(paths) - len(todo)
    if summary.skipped:
        log(f"Skipping {summary.skipped} already processed frames")
    # shard on CLIP POSITIONS (i % world over the unfiltered, sorted list), then drop what is done: which rank owns a
    # frame, and FrameOutput.index, never depend on how far an earlier run got
    todo_set = set(todo)
    mine = [(i, paths[i]) for i in video.shard_frames(len(paths), rank, world) if paths[i] in todo_set]
    for frame in FrameLoader(mine, downscale_factor, decode_threads, load_fn=load_fn):
        ok = False
        try:
            if frame.image is None:
                raise RuntimeError(frame.error)
            pred = model.infer(frame.image, f_px=frame.f_px)
            depth, focal = pred["depth"], float(pred["focallength_px"])
            out = FrameOutput(frame.index, frame.path, depth, focal, frame.image)
            if unproject:
                h, w = depth.shape
                import torch

                rgb = torch.from_numpy(frame.image).to(depth.device)
                out.points, out.valid_mask, out.colors = video.depth_to_3d(model, depth, focal, w, h, rgb=rgb)
            ok = consumer(out) is not False
        except Exception as e:  # noqa: BLE001 -- per-frame isolation, as in the reference
            log(f"Error processing {frame.path}: {e}")
        progress.mark(frame.path, ok)
        if ok:
            summary.processed += 1
        else:
            summary.failed.append(os.path.basename(frame.path))
    progress.save()
    if world > 1:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.barrier()
    if rank == 0:
        progress.merge_shards()
    summary.seconds = time.time() - t0
    return summary
