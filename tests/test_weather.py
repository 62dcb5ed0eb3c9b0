"""§8 f3: the rain sequence of a reference run.  `Weather::make_pattern` (src/weather.rs:19-50) draws from
`StdRng::from_seed([0; 32])`, which in the pinned rand 0.5.3 is HC-128.  The crate is not under /root/reference, so the
cipher is restated from its specification and pinned on the specification's all-zero key / IV keystream (the same 16
words rand's own `test_hc128_true_values_a` holds); the Python and C++ hosts must agree word for word and day for day."""
import os
import subprocess

import numpy as np
import pytest

from motivate_b200 import weather

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "motivate_b200", "motivate_b200_run")

# HC-128 specification, test vector 1: key = 0, IV = 0, first 512 bits of keystream as 32-bit words
HC128_ZERO_KEYSTREAM = [0x73150082, 0x3bfd03a0, 0xfb2fd77f, 0xaa63af0e, 0xde122fc6, 0xa7dc29b6, 0x62a68527, 0x8b75ec68,
                        0x9036db1e, 0x81896005, 0x00ade078, 0x491fbf9a, 0x1cdc3013, 0x6c3d6e24, 0x90f664b2, 0x9cd57102]


def test_hc128_known_answer():
    rng = weather.Hc128()
    assert [rng.next_u32() for _ in range(16)] == HC128_ZERO_KEYSTREAM


def test_std_rng_word_order_and_float_conversion():
    rng = weather.StdRng(bytes(32))
    assert rng.next_u64() == 0x3bfd03a073150082                        # low word first (rand_core BlockRng::next_u64)
    u = weather.StdRng(bytes(32)).gen_f64()
    assert u == (0x3bfd03a073150082 >> 11) / 2.0 ** 53 and 0.0 <= u < 1.0
    assert abs(u - 0.23433) < 1e-5


def test_make_pattern_is_the_chain_of_the_draws():
    n = 1825                                                           # 5 years, parameters.yaml of the reference
    w = weather.make_pattern(n)
    rng = weather.StdRng(bytes(32))
    u = [rng.gen_f64() for _ in range(n)]
    assert w[0] == (0 if u[0] > 0.14 else 1)
    for i in range(1, n):
        assert w[i] == (0 if u[i] < (0.886 if w[i - 1] == 0 else 0.699) else 1)
    assert 0.08 < w.mean() < 0.20                                      # stationary rain share 0.114 / (0.114 + 0.699) = 0.14
    assert len(weather.make_pattern(0)) == 0
    # another matrix / first-day chance goes through the same draws
    always_bad = weather.make_pattern(50, {0: 0.0, 1: 0.0}, chance_of_rain=1.0)
    assert always_bad.all()


@pytest.mark.skipif(not os.path.exists(RUNNER), reason="C++ host not built")
def test_cpp_host_agrees():
    words = subprocess.run([RUNNER, "--keystream", "2100"], capture_output=True, text=True, check=True).stdout.split()
    rng = weather.Hc128()
    assert [int(x, 16) for x in words] == [rng.next_u32() for _ in range(2100)]          # crosses both table halves twice
    assert [int(x, 16) for x in words[:16]] == HC128_ZERO_KEYSTREAM
    days = subprocess.run([RUNNER, "--weather", "1825"], capture_output=True, text=True, check=True).stdout.strip()
    assert days == "".join(str(int(x)) for x in weather.make_pattern(1825))
