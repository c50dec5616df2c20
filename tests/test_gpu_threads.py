"""Several contexts of one process working at the same time (one host thread each, as bench.py's lanes and the two
contexts of unconvertDWfile do): every thread must get exactly the bytes a lone context gets.  Blocks are large enough
for the library's copy gate (copies of 8 MiB and more take turns, per device and direction)."""
import threading

import pytest

import gpuutil as G
import oracle as O

pytestmark = pytest.mark.gpu


def test_contexts_on_threads_do_not_disturb_each_other():
    from zdw_b200 import Context

    work = []
    for name in ("movie_tickets", "analytics-hits"):
        tsv = O.golden(f"{name}.sql")
        sch = O.parse_desc(O.golden(f"{name}.desc.sql"))
        _, block = G.split_header(O.golden_to_v11(O.golden(f"{name}.zdw")))  # the reference's own encoding
        work.append((name, sch, tsv, block))
    assert len(work[0][2]) >= (8 << 20) and len(work[0][3]) >= (8 << 20)  # both directions pass the gate's threshold

    errors = []
    start = threading.Barrier(3)

    def lane(k: int):
        try:
            with Context(0) as ctx:
                start.wait(timeout=60)
                for it in range(4):
                    name, sch, tsv, block = work[(k + it) % len(work)]
                    blk = ctx.encode_block(sch.types, tsv)
                    if blk.data != block:
                        errors.append(f"thread {k} pass {it} {name}: encode {G.first_diff(blk.data, block)}")
                        return
                    dec = ctx.decode_block(sch.types, block)
                    if dec.tsv != tsv:
                        errors.append(f"thread {k} pass {it} {name}: decode {G.first_diff(dec.tsv, tsv)}")
                        return
        except Exception as ex:  # noqa: BLE001
            errors.append(f"thread {k}: {ex!r}")

    threads = [threading.Thread(target=lane, args=(k,)) for k in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in threads), "a lane did not finish"
    assert not errors, errors
