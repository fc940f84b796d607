"""CPU tests of host-side helpers that need no GPU."""
import os

from gpu_image_processing_b200 import affinity


def test_cpulist_parsing():
    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()
    assert affinity._parse_cpulist("5") == {5}


def test_binding_is_best_effort_without_a_gpu():
    before = os.sched_getaffinity(0)
    report = affinity.bind_to_device_numa(0)
    assert isinstance(report, dict) and "bound" in report
    if not report["bound"]:
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)


def test_binding_uses_only_allowed_local_cpus(monkeypatch):
    before = os.sched_getaffinity(0)
    some = set(sorted(before)[: max(1, len(before) // 2)])
    monkeypatch.setattr(affinity, "device_local_cpus", lambda i: some | {10 ** 6})
    try:
        report = affinity.bind_to_device_numa(0)
        if some == before:
            assert not report["bound"]
        else:
            assert report["bound"] and os.sched_getaffinity(0) == some
    finally:
        os.sched_setaffinity(0, before)
