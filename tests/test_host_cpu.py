"""CPU: the host mirror's droplet queue against the reference's own queue tests
(server/libs/queue/queue_test.go:54-176): overwrite order, Gets batching, release-on-overwrite
products, flush indicator."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H(built):
    from deepflow_b200 import build
    path = build.build_host()
    try:
        L = C.CDLL(path)
    except OSError as e:  # libdfgpu.so needs libcudart at load time
        pytest.skip(f"libdfhost.so not loadable here: {e}")
    L.dfhost_queue_new.restype = C.c_void_p
    L.dfhost_queue_new.argtypes = [C.c_int, C.c_int, C.c_int]
    L.dfhost_queue_free.argtypes = [C.c_void_p]
    L.dfhost_queue_put.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
    L.dfhost_queue_get.restype = C.c_longlong
    L.dfhost_queue_get.argtypes = [C.c_void_p]
    L.dfhost_queue_gets.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
    L.dfhost_queue_len.argtypes = [C.c_void_p]
    L.dfhost_queue_size.argtypes = [C.c_void_p]
    L.dfhost_queue_released.restype = C.c_longlong
    L.dfhost_queue_released.argtypes = [C.c_void_p, C.c_int]
    return L


class Q:
    def __init__(self, L, size, release=False, flush_ms=0):
        self.L, self.h = L, L.dfhost_queue_new(size, 1 if release else 0, flush_ms)

    def put(self, *vals):
        a = (C.c_longlong * len(vals))(*vals)
        return self.L.dfhost_queue_put(self.h, a, len(vals))

    def get(self):
        return self.L.dfhost_queue_get(self.h)

    def gets(self, cap):
        a = (C.c_longlong * max(cap, 1))()
        n = self.L.dfhost_queue_gets(self.h, a, cap)
        return list(a)[:n]

    def released(self, reset=False):
        return self.L.dfhost_queue_released(self.h, 1 if reset else 0)

    def close(self):
        self.L.dfhost_queue_free(self.h)


def test_queue_overwrite_and_size(H):
    q = Q(H, 2)  # TestQueueOverWrite / TestQueueOverSize
    q.put(10086, 10087)
    q.put(10088)
    assert q.get() == 10087 and q.get() == 10088
    q.close()
    q = Q(H, 3)  # TestQueueSize: capacity rounds up to 4
    assert H.dfhost_queue_size(q.h) == 4
    q.put(10086, 10087, 10088, 10089)
    q.gets(3)
    assert q.get() == 10089
    assert q.put(1, 2, 3, 4, 5) == -1  # OverflowError
    q.close()


def test_queue_gets(H):
    q = Q(H, 2)  # TestQueueGets
    q.put(10086)
    q.put(10087, 10088)
    assert q.gets(2) == [10087, 10088]
    q.put(10086, 10087)
    assert q.gets(100) == [10086, 10087]
    q.close()
    q = Q(H, 1)  # TestQueueEmptyGets
    q.put(10086)
    assert q.gets(0) == []
    assert q.gets(1) == [10086]
    q.close()
    q = Q(H, 5)  # TestQueueTwiceGets
    q.put(10086, 10087, 10088, 10089, 10090)
    assert q.gets(2) == [10086, 10087]
    assert q.gets(3) == [10088, 10089, 10090]
    q.close()


def test_queue_release_products(H):
    q = Q(H, 4, release=True)  # TestReleased
    q.put(2, 3, 5, 7)
    q.put(11)
    assert q.released() == 2
    q.put(13, 17)
    q.released(reset=True)
    q.put(19, 23)
    assert q.released() == 77
    q.put(29, 31)
    q.released(reset=True)
    q.put(37)
    assert q.released() == 19
    q.close()
    q = Q(H, 4, release=True)  # TestPartialRelease
    q.put(1, 2, 3)
    q.put(5, 7)
    assert q.released() == 1
    assert q.get() == 2
    q.put(11, 13)
    assert q.released() == 3
    q.get()
    q.get()
    q.released(reset=True)
    q.put(17, 19, 23)
    assert q.released() == 11
    q.close()


def test_queue_flush_indicator(H):
    q = Q(H, 1, flush_ms=1)  # TestFlushIndicator: a nil item arrives by itself
    assert q.get() == 0
    q.close()
