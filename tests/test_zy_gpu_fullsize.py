"""The BASELINE.json configurations at their FULL sizes against the oracle (>= 10 steps where the oracle finishes in a
minute or two on the GPU box's host cores; 4 steps at the 4096^2 headline size), fields and diagnostics to 1e-10.
The file name sorts these after the scaled cases of test_gpu_parity.py (largest oracle runs of the suite)."""
import pytest

import test_gpu_parity as T

pytestmark = pytest.mark.gpu
gpu, O = T.gpu, T.O


def test_config2_slqg_1024(gpu, O):
    """configs[1]: SingleLayerQG beta-plane stochastic 1024^2 FilteredRK4."""
    T.config2_slqg(gpu, O, 1024, 12, 4)


def test_config3_mqg_2layer_1024(gpu, O):
    """configs[2]: MultiLayerQG 2-layer 1024^2 FilteredRK4 (examples/multilayerqg_2layer.jl:32-63)."""
    T.config3_mqg(gpu, O, 2, 1024, 10, 5)


@pytest.mark.timeout(600)
def test_config3_mqg_5layer_2048(gpu, O):
    """configs[2], 5 layers at 2048^2."""
    T.config3_mqg(gpu, O, 5, 2048, 10, 5)


@pytest.mark.timeout(600)
def test_config4_sqg_4096(gpu, O):
    """configs[3], the benchmarked problem itself: SurfaceQG 4096^2 ETDRK4, elliptical vortex, 4 steps vs the oracle."""
    T.config4_sqg(gpu, O, "ETDRK4", 4096, 4, 2)
