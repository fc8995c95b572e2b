"""Host <-> device feed helpers (mmrseg_b200.data): ResultReader hands every step's value to the host, in order,
one step late, without a full synchronisation per step (the reference reads `loss.item()` after every step,
SU/ModelTraining.py:619-622)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_result_reader_returns_every_value_in_order():
    from mmrseg_b200.data import ResultReader
    reader = ResultReader(lag=1)
    acc = torch.zeros((), device="cuda")
    got = []
    for i in range(7):
        acc = acc + (i + 1)                 # a fresh tensor per step, like a loss
        out = reader.push(acc)
        assert len(out) == (1 if i >= 1 else 0)
        got += out
    got += reader.flush()
    assert [float(v) for v in got] == [1.0, 3.0, 6.0, 10.0, 15.0, 21.0, 28.0]
    assert reader.flush() == []


def test_result_reader_reused_buffer_and_matrix_results():
    from mmrseg_b200.data import ResultReader
    reader = ResultReader(lag=2)
    buf = torch.zeros((3, 3), dtype=torch.int64, device="cuda")   # a result buffer the step overwrites in place
    got = []
    for i in range(5):
        buf.fill_(i)
        got += reader.push(buf)
    got += reader.flush()
    assert [int(v[1, 1]) for v in got] == [0, 1, 2, 3, 4] and all(v.shape == (3, 3) for v in got)
