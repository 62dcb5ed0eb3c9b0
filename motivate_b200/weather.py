"""The rain sequence a run of the reference would use.

`Weather::make_pattern` (src/weather.rs:19-50) draws `rng.gen::<f64>()` from `StdRng::from_seed([0; 32])`.  The crates
behind that call are not under /root/reference; Cargo.lock pins rand 0.5.3 / rand_core 0.2.1, where
  * `StdRng` is `Hc128Rng`: the HC-128 stream cipher (Wu, eSTREAM portfolio) keyed with seed bytes 0-15, IV = seed
    bytes 16-31 (little-endian words), its keystream words handed out in blocks of 16 (`BlockRng`);
  * `next_u64` takes two consecutive words, low word first;
  * `gen::<f64>()` is `(next_u64() >> 11) as f64 * 2^-53` (53 random bits, [0, 1)).
HC-128 is restated here from its published specification; tests/test_weather.py pins it on the specification's
all-zero key / IV keystream.  Host-side set-up code: a few thousand draws per run, so plain Python integers do.
"""
import numpy as np

M32 = 0xFFFFFFFF
GOOD, BAD = 0, 1


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & M32


def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & M32


class Hc128:
    """HC-128 keystream generator: two 512-word tables P, Q updated in turn, 512 steps each."""

    def __init__(self, key_words=(0, 0, 0, 0), iv_words=(0, 0, 0, 0)):
        w = list(key_words) * 2 + list(iv_words) * 2
        for i in range(16, 1280):
            f2 = _rotr(w[i - 2], 17) ^ _rotr(w[i - 2], 19) ^ (w[i - 2] >> 10)
            f1 = _rotr(w[i - 15], 7) ^ _rotr(w[i - 15], 18) ^ (w[i - 15] >> 3)
            w.append((f2 + w[i - 7] + f1 + w[i - 16] + i) & M32)
        self.p, self.q, self.i = w[256:768], w[768:1280], 0
        for _ in range(1024):                       # 1024 steps whose output is folded back into the tables
            j = self.i & 511
            t = self.p if self.i & 512 == 0 else self.q
            s = self._step()
            t[j] = s
        self.i = 0

    def _step(self):
        i = self.i
        j = i & 511
        self.i = (i + 1) & 1023
        if i & 512 == 0:
            t, o = self.p, self.q
            g = ((_rotr(t[(j - 3) & 511], 10) ^ _rotr(t[(j - 511) & 511], 23)) + _rotr(t[(j - 10) & 511], 8)) & M32
        else:
            t, o = self.q, self.p
            g = ((_rotl(t[(j - 3) & 511], 10) ^ _rotl(t[(j - 511) & 511], 23)) + _rotl(t[(j - 10) & 511], 8)) & M32
        t[j] = (t[j] + g) & M32
        x = t[(j - 12) & 511]
        h = (o[x & 0xFF] + o[256 + ((x >> 16) & 0xFF)]) & M32
        return h ^ t[j]

    def next_u32(self):
        return self._step()


class StdRng:
    """rand 0.5.3 `StdRng::from_seed(seed)` as far as this path uses it: `gen::<f64>()` only."""

    def __init__(self, seed=bytes(32)):
        if len(seed) != 32:
            raise ValueError("StdRng seed is 32 bytes")
        words = [int.from_bytes(seed[4 * k:4 * k + 4], "little") for k in range(8)]
        self._core = Hc128(words[:4], words[4:])

    def next_u64(self):
        lo = self._core.next_u32()
        return lo | (self._core.next_u32() << 32)

    def gen_f64(self):
        return float(self.next_u64() >> 11) * (1.0 / 9007199254740992.0)


#: the matrix `main` passes (src/main.rs:73-85): P(next is Good | today)
TRANSITION_TO_GOOD = {GOOD: 0.886, BAD: 0.699}
CHANCE_OF_RAIN = 0.14


def make_pattern(days, transition_to_good=None, chance_of_rain=CHANCE_OF_RAIN, seed=bytes(32)):
    """`Weather::make_pattern` (src/weather.rs:19-50): day 0 is Good when the draw exceeds `chance_of_rain`; day i is
    Good when the draw is below P(Good | weather of day i-1).  Returns uint8[days], 0 Good / 1 Bad."""
    tg = TRANSITION_TO_GOOD if transition_to_good is None else transition_to_good
    rng = StdRng(seed)
    w = np.zeros(days, np.uint8)
    if days == 0:
        return w
    w[0] = GOOD if rng.gen_f64() > chance_of_rain else BAD
    for i in range(1, days):
        w[i] = GOOD if rng.gen_f64() < tg[int(w[i - 1])] else BAD
    return w
