"""Section timer with the device's view: the B200 counterpart of the reference's nested wall-clock profiler
(/root/reference/ns/lib/profiler.py:4-52; switched on by utils/train_dataset.py:52, wraps the model inference at :97).

    Profiler.enabled = True
    with Profiler('label'):
        ...                      # nested sections are reported as an indented tree when the outermost one closes

Every section reports three things: host wall-clock seconds (what the reference prints), the time between two CUDA
events recorded on the current stream at its entry and exit (host launch overhead excluded), and how many kernels of
libmlamg_b200.so were launched inside it.  Output lines keep the reference's shape `[label] 0.01234s`, the device part
is appended in parentheses.  Like the reference's, a section swallows an exception raised inside it (its callers sit in
the try / except of the fitness loop, utils/train_dataset.py:95-103), and a disabled profiler is a silent no-op.
"""
import time

import torch

import mlamg


class _Mark:
    """host clock, an optional CUDA event on the current stream and the library's launch counter, read together"""
    __slots__ = ("wall", "event", "launches")

    def __init__(self):
        self.event = None
        if torch.cuda.is_available():
            self.event = torch.cuda.Event(enable_timing=True)
            self.event.record()
        self.launches = mlamg.launch_count()
        self.wall = time.time()


class Profiler:
    enabled = False          # class-level switch, as in the reference
    current = None           # innermost open section
    tab_width = 2

    def __init__(self, section_name):
        self.section_name = section_name
        self.children = []
        self.parent = None
        self._begin = self._end = None

    # ---- context manager ------------------------------------------------------------------------
    def __enter__(self):
        if Profiler.enabled:
            self.parent, Profiler.current = Profiler.current, self
            if self.parent is not None:
                self.parent.children.append(self)
            self._begin = _Mark()
            return self

    def __exit__(self, exc_type, exc, tb):
        if Profiler.enabled and self._begin is not None:
            self._end = _Mark()
            Profiler.current = self.parent
            if self.parent is None:
                if self._end.event is not None:
                    self._end.event.synchronize()            # the last event of the tree: every other one is complete
                for depth, node in self._walk(0):
                    print(' ' * (depth * Profiler.tab_width) + node._line())
        return True

    # ---- results --------------------------------------------------------------------------------
    @property
    def running_time(self):
        return self._end.wall - self._begin.wall

    @property
    def device_time(self):
        if self._begin.event is None or self._end.event is None:
            return None
        return self._begin.event.elapsed_time(self._end.event) * 1e-3

    @property
    def kernel_launches(self):
        return self._end.launches - self._begin.launches

    def _walk(self, depth):
        yield depth, self
        for child in self.children:
            if child._end is not None:
                yield from child._walk(depth + 1)

    def _line(self):
        text = f'[{self.section_name}] {self.running_time:.5f}s'
        dev = self.device_time
        if dev is not None:
            text += f' (device {dev:.5f}s, {self.kernel_launches} kernels)'
        return text
