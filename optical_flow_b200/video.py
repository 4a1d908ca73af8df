"""Host-side frame access for the two entry-point scripts.  No arithmetic of the hot
path lives here: decoding and JPEG encoding stay with cv2 exactly as in the reference; this module only removes
avoidable host latency around them.

FrameReader   the reference seeks before EVERY frame (`vid.set(CAP_PROP_POS_FRAMES, pos); vid.read()`,
              /root/reference/optical_flow.py:36-37, visualize_optical_flow.py:23-24).  A seek restarts decoding at the
              previous key frame.  When the requested frame lies a few frames ahead of the decoder's position the same
              frame is reached by grabbing forward; anything else (first access, backward jump, long jump) seeks as the
              reference does.  cv2 truncates a float position to an integer frame index, and so does this.
JpegWriter    `cv2.imwrite` calls (visualize_optical_flow.py:57-60) handed to a small thread pool (cv2 releases the GIL
              while encoding), so encoding overlaps decoding and the GPU; same encoder, same bytes.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import cv2


class FrameReader:
    def __init__(self, vid, max_forward=48):
        self.vid = vid
        self.max_forward = int(max_forward)
        self._next = None          # index of the frame the next vid.read() would return, when known
        self.seeks = 0
        self.grabs = 0

    def read_at(self, pos):
        """(ok, frame) exactly as `vid.set(cv2.CAP_PROP_POS_FRAMES, pos); vid.read()` returns them."""
        target = int(pos)
        ahead = None if self._next is None else target - self._next
        forward = ahead is not None and 0 <= ahead <= self.max_forward
        if forward:
            for _ in range(ahead):
                self.grabs += 1
                if not self.vid.grab():
                    self._next = None
                    return False, None
            # cv2 resolves a seek from the stream's timestamps (dts_to_frame_number); on VFR / drop-frame streams counting grabs
            # can land elsewhere.  The decoder reports the index of the frame it will return next: if that is not the
            # target, do what the reference does.
            if int(self.vid.get(cv2.CAP_PROP_POS_FRAMES)) != target:
                forward = False
        if not forward:
            self.seeks += 1
            self.vid.set(cv2.CAP_PROP_POS_FRAMES, pos)
        ok, frame = self.vid.read()
        self._next = target + 1 if ok else None
        return ok, frame


class JpegWriter:
    def __init__(self, workers=None):
        n = workers or max(1, min(8, (os.cpu_count() or 2) - 1))
        self._pool = ThreadPoolExecutor(max_workers=n)
        self._pending = []

    def imwrite(self, path, image):
        """Queues cv2.imwrite(path, image); the image must not be modified until close()."""
        self._pending.append(self._pool.submit(cv2.imwrite, path, image))

    def close(self):
        ok = all(f.result() for f in self._pending)
        self._pending = []
        self._pool.shutdown(wait=True)
        return ok

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
