"""TaichiQueue compatibility shim (reference: taichi_splatting/taichi_queue.py:34-90).

The reference funnels every Taichi launch through one executor because the Taichi runtime is not
thread-safe.  This package has no Taichi runtime: kernels are plain CUDA launches on the caller's current
torch stream and the library is re-entrant, so the queue is a no-op that keeps the call sites working.
"""
from concurrent.futures import Future


class TaichiQueue:
  _initialised = False

  @classmethod
  def init(cls, *args, threaded=False, **kwargs) -> None:
    cls._initialised = True

  @classmethod
  def stop(cls) -> None:
    cls._initialised = False

  @staticmethod
  def thread_id():
    return None

  @staticmethod
  def run_sync(func, *args, **kwargs):
    args = [a.result() if isinstance(a, Future) else a for a in args]
    return func(*args, **kwargs)

  @staticmethod
  def run_async(func, *args, **kwargs) -> Future:
    future = Future()
    future.set_result(TaichiQueue.run_sync(func, *args, **kwargs))
    return future


class _QueueContext:
  def __init__(self, *args, **kwargs):
    self.args, self.kwargs = args, kwargs

  def __enter__(self):
    TaichiQueue.init(*self.args, **self.kwargs)

  def __exit__(self, exc_type, exc_value, traceback):
    TaichiQueue.stop()


def taichi_queue(*args, **kwargs):
  return _QueueContext(*args, **kwargs)


def queued(fn):
  return fn
