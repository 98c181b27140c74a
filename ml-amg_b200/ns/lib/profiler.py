"""Mirror of the reference's nested section profiler (/root/reference/ns/lib/profiler.py:4-52; switched on by
utils/train_dataset.py:52, wraps the model inference at :97) with the device's view added: besides the wall-clock time
of a section, the time between two CUDA events recorded on the current stream at its entry and exit, and the number
of libmlamg_b200 kernels launched inside it.  Same usage and output shape:

    Profiler.enabled = True
    with Profiler('label'):
        ...                      # nested `with` blocks print hierarchically when the root section closes

As in the reference, `__exit__` returns True: an exception raised inside a section is swallowed (its callers rely on the
surrounding try / except of the fitness loop, utils/train_dataset.py:95-103).
"""
import time

import torch

import mlamg


class Profiler:
    enabled = False
    current = None
    tab_width = 2

    def __init__(self, section_name):
        self.section_name = section_name
        self.children = []

    def __enter__(self):
        if not Profiler.enabled:
            return
        self.parent = Profiler.current
        if self.parent is not None:
            self.parent.children.append(self)
        Profiler.current = self
        self._events = None
        if torch.cuda.is_available():
            self._events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._events[0].record()
        self._launches = mlamg.launch_count()
        self.start_time = time.time()
        return self

    def _resolve(self):
        self.device_time = None
        if self._events is not None:
            self.device_time = self._events[0].elapsed_time(self._events[1]) * 1e-3
        for child in self.children:
            child._resolve()

    def print_recursive(self, level):
        dev = '' if self.device_time is None else f' (device {self.device_time:.5f}s, {self.kernel_launches} kernels)'
        print(((level * Profiler.tab_width) * ' ') + f'[{self.section_name}] {self.running_time:.5f}s' + dev)
        for child in self.children:
            child.print_recursive(level + 1)

    def __exit__(self, type, value, tb):
        if not Profiler.enabled:
            return True
        if self._events is not None:
            self._events[1].record()
        self.running_time = time.time() - self.start_time
        self.kernel_launches = mlamg.launch_count() - self._launches
        if self.parent is None:
            if self._events is not None:
                torch.cuda.synchronize()
            self._resolve()
            self.print_recursive(0)
        Profiler.current = self.parent
        return True
