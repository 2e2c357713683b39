"""Slab-sharded imfilter: world_size-2/3 process groups.

CPU suite (gloo, host tensors): the per-slab compute is the oracle library, so what is tested is sharded.py's host
logic — partition, neighbours (incl. the circular wrap), message order, halo sizes, global-coordinate borders.
GPU suite (-m gpu): the same worker with the product library on cuda:0, both halo transports ("p2p" = CUDA IPC peer
pointers read by the fused kernel, "staged" = copy-engine staging overlapped with the kernel behind flag bytes, "sendrecv" =
exchanged halo buffers)."""
import pytest

from sharded_worker import launch, run_mapwindow


def _check(results, world):
    assert len(results) == world
    for rank, failures in results:
        assert failures == [], (rank, failures)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_host_logic_gloo_cpu(world):
    _check(launch(world, use_device=False, modes=["sendrecv"]), world)


def test_slab_bounds_and_halo_extent(ifb):
    from importlib import import_module
    sh = import_module("imagefiltering_jl_b200.sharded")
    imf = import_module("imagefiltering_jl_b200.imfilter")
    assert [sh.slab_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert [sh.slab_bounds(1024, 8, r) for r in range(8)] == [(128 * r, 128) for r in range(8)]
    st = imf.build_stages(ifb.KernelFactors.gaussian((4, 4, 4)), 3)
    assert sh.halo_extent(st, 3) == (8, 8)
    st = imf.build_stages(ifb.KernelFactors.gaussian((4, 4, 0)), 3)
    assert sh.halo_extent(st, 3) == (0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_device_two_ranks_one_gpu(world):
    _check(launch(world, use_device=True, modes=["driver", "p2p", "staged", "sendrecv"]), world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_driver_xy_filtered_exchange(world):
    """B2F_SHARD_XY=1: the driver exchanges xy-filtered boundary planes (b2f_imfilter_slab_xy) instead of raw halos"""
    _check(launch(world, use_device=True, modes=["driver"], env={"B2F_SHARD_XY": "1"}), world)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_mapwindow_host_logic_gloo_cpu(world):
    """ShardedMapwindow (one huge volume): partition, halo planes of the window, global faces"""
    _check(launch(world, use_device=False, modes=[], target=run_mapwindow), world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_mapwindow_device(world):
    _check(launch(world, use_device=True, modes=[], target=run_mapwindow), world)
