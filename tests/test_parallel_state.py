"""Group-layout ground truths (role of reference test/unit_test/parallel_layers/test_parallel_state.py)."""
import pytest

from neuronx_distributed_b200.parallel_layers import parallel_state as ps
from neuronx_distributed_b200.parallel_layers.parallel_state import RankMesh, arrange_kv_groups


def test_mesh_tp_dp_pp_layout():
    m = RankMesh(world_size=16, tp=2, pp=2, cp=1, ep=1)
    assert m.dp == 4
    assert m.tp_groups()[:3] == [[0, 1], [2, 3], [4, 5]]
    assert m.dp_groups()[0] == [0, 2, 4, 6] and m.dp_groups()[1] == [1, 3, 5, 7]
    assert m.pp_groups()[0] == [0, 8] and m.pp_groups()[-1] == [7, 15]
    assert m.coords(11) == (1, 1, 0, 1)


def test_mesh_128_ranks_tp8_pp4():
    m = RankMesh(world_size=128, tp=8, pp=4, cp=1, ep=1)
    assert m.dp == 4
    assert m.tp_groups()[0] == list(range(8))
    assert m.tp_groups()[-1] == list(range(120, 128))
    assert m.dp_groups()[0] == [0, 8, 16, 24]
    assert m.pp_groups()[0] == [0, 32, 64, 96]
    assert len(m.tp_groups()) == 16 and len(m.dp_groups()) == 32 and len(m.pp_groups()) == 32


def test_mesh_context_parallel_and_zero1_groups():
    m = RankMesh(world_size=16, tp=2, pp=1, cp=2, ep=1)
    assert m.dp == 4
    assert m.cp_groups()[0] == [0, 2] and m.cp_groups()[1] == [1, 3]
    assert m.dp_groups()[0] == [0, 4, 8, 12]
    # ZeRO-1 shards over DP x CP
    z = m.zero1_sharding_groups()
    assert z[0] == [0, 2, 4, 6, 8, 10, 12, 14] and z[1] == [1, 3, 5, 7, 9, 11, 13, 15]
    assert m.cp_ring_pairs(0) == [(0, 2), (2, 0)]


def test_mesh_expert_parallel():
    m = RankMesh(world_size=16, tp=2, pp=1, cp=1, ep=4)
    assert m.dp == 8 and m.dp_exp == 2
    assert m.ep_model_groups()[0] == [0, 2, 4, 6]
    assert m.ep_data_groups()[0] == [0, 8]
    assert m.coords_exp(10) == (0, 1, 1, 0)


def test_invalid_degrees():
    with pytest.raises(RuntimeError):
        RankMesh(world_size=6, tp=4, pp=1, cp=1, ep=1)
    with pytest.raises(RuntimeError):
        RankMesh(world_size=8, tp=2, pp=1, cp=1, ep=3)


def test_kv_groups():
    assert arrange_kv_groups(1, 4, 2, False) == [[0, 2], [1, 3]]
    assert arrange_kv_groups(1, 4, 2, True) == [[0, 1], [2, 3]]
    assert arrange_kv_groups(2, 4, 2, False) == [[0, 2], [1, 3], [4, 6], [5, 7]]
    assert arrange_kv_groups(1, 8, 4, False, adjacent_replication=True) == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_mesh_only_api():
    m = ps.initialize_model_parallel(4, 2, mesh_only=True, world_size=32)
    assert (m.tp, m.pp, m.dp) == (4, 2, 4)


def test_experts_for_rank():
    assert ps.get_experts_for_expert_parallel_rank(1, 8, 4) == [2, 3]
    assert ps.get_experts_for_expert_parallel_rank(0, 4, 2, [[3, 1], [0, 2]]) == [3, 1]
