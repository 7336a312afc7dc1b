"""Observability formats the reference's analysis scripts read: the TIME_BENCH label table
(svg/timer.py) and the density JSONL schema (svg/models/hyvideo/attention.py:786-802, svg/utils/density.py)."""
import json

import pytest
import torch


def test_timer_table_format_matches_reference_layout(monkeypatch):
    from svgb200 import timer

    monkeypatch.setattr(timer, "operator_log_data", {"Level 3 - sample mse": 1234.5, "Level 2 - qkv": 20.0})
    monkeypatch.setattr(timer, "CLEAR_LOG_DATA", False)
    lines = timer.format_operator_log_data().split("\n")
    assert lines[0].startswith("Level 2 - qkv        :") and lines[0].endswith(" s")       # sorted, padded, seconds
    assert lines[1] == "Level 3 - sample mse :        1.23 s"
    monkeypatch.setattr(timer, "CLEAR_LOG_DATA", True)
    assert timer.format_operator_log_data().split("\n")[1] == "Level 3 - sample mse :     1234.50 ms"
    # disabled by default: the context manager records nothing and never synchronises
    assert timer.ENABLE_LOGGING is False
    with timer.time_logging_decorator("x"):
        pass
    assert "x" not in timer.operator_log_data


@pytest.mark.gpu
def test_density_jsonl_schema(cuda, tmp_path):
    from svgb200.models import wan

    F, P, H, D = 4, 64, 2, 64
    S = F * P
    g = torch.Generator(device=cuda).manual_seed(0)
    q, k, v = (torch.randn(1, H, S, D, device=cuda, generator=g).bfloat16() for _ in range(3))
    log = tmp_path / "density.jsonl"
    sap = wan.WanSAPCore(F, P, num_q_centroids=4, num_k_centroids=8, top_p_kmeans=0.9, min_kc_ratio=0.1,
                         kmeans_iter_init=3, kmeans_iter_step=1, first_times_fp=1000)
    sap.logging_file = str(log)
    sap.attention_core_logic(q, k, v, torch.tensor([500]), layer_idx=7)
    sap.attention_core_logic(q, k, v, torch.tensor([400]), layer_idx=7)
    rows = [json.loads(x) for x in log.read_text().splitlines()]
    assert len(rows) == 2 and set(rows[0]) == {"timestep", "layer", "avg_density", "density"}
    assert rows[0]["timestep"] == 500 and rows[1]["timestep"] == 400 and rows[0]["layer"] == 7
    dens = torch.tensor(rows[0]["density"])
    assert dens.shape == (1, H) and 0 < rows[0]["avg_density"] <= 1
    assert abs(dens.mean().item() - rows[0]["avg_density"]) < 1e-6
