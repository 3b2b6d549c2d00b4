"""The pin of the CPU restatement against the UNMODIFIED reference (tests/test_oracle_vs_reference.py), once more under the
`gpu` marker so that the driver's GPU-box run (`-m gpu`, which deselects the unmarked tests) executes it there as well:
oracle/_ref/libtake_ref.so travels with the snapshot.  Skipped where that library is absent."""
import pytest

import test_oracle_vs_reference as pin

pytestmark = pytest.mark.gpu

test_rng_preload_trick = pin.test_rng_preload_trick
test_scene_generators_match_reference_parser = pin.test_scene_generators_match_reference_parser
test_port_equals_reference = pin.test_port_equals_reference
test_default_max_depth_50 = pin.test_default_max_depth_50
