"""rc_pyref.py -- second, independent CPU restatement (TEST INFRASTRUCTURE ONLY).

A type-by-type Python transliteration of the reference crate's semantics using
Python integers masked to 64/32 bits.  It exists only to cross-check
oracle/rc_oracle.c on small inputs (two restatements written separately must
agree byte for byte) -- see the "PARITY UNPINNED" note in rc_oracle.h.
Only tests/ may import this module.

Citations are path:line under /root/reference/.
"""
from collections import deque

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


class RangeCoderError(Exception):
    """src/error.rs:3-13"""


class LowerBoundOverflow(RangeCoderError):
    pass


class UpperBoundOverflow(RangeCoderError):
    pass


class RangeCoder:
    """src/range_coder.rs:7-147"""

    TOP8 = 1 << (64 - 8)  # :23
    TOP16 = 1 << (64 - 16)  # :24

    def __init__(self):  # :13-20
        self.lower_bound = 0
        self.range = M64

    def range_par_total(self, total_freq):  # :38-40
        if total_freq == 0:
            raise ZeroDivisionError("attempt to divide by zero")
        return self.range // total_freq

    def upper_bound(self):  # :138-146
        u = self.lower_bound + self.range
        if u > M64:
            raise UpperBoundOverflow(self.lower_bound, self.range)
        return u

    def left_shift(self):  # :95-100
        tmp = (self.lower_bound >> 56) & 0xFF
        self.range = (self.range << 8) & M64
        self.lower_bound = (self.lower_bound << 8) & M64
        return tmp

    def no_carry_expansion(self):  # :110-116
        if (self.lower_bound ^ self.upper_bound()) < self.TOP8:
            return self.left_shift()
        return None

    def range_reduction_expansion(self):  # :126-135
        if self.range < self.TOP16:
            self.range = (~self.lower_bound) & (self.TOP16 - 1)
            return self.left_shift()
        return None

    def param_update(self, c_freq, cum_freq, total_freq, limit=64):  # :53-92
        out = deque()
        rpt = self.range_par_total(total_freq)
        self.range = (rpt * c_freq) & M64
        add = (rpt * cum_freq) & M64
        nl = self.lower_bound + add
        if nl > M64:
            raise LowerBoundOverflow(self.lower_bound, add, self.range)
        self.lower_bound = nl
        while True:
            b = self.no_carry_expansion()
            if b is None:
                break
            out.append(b)
            if len(out) > limit:  # the reference would never return (range == 0)
                raise RuntimeError("loop 1 does not terminate (zero frequency symbol)")
        while True:
            b = self.range_reduction_expansion()
            if b is None:
                break
            out.append(b)
        return out


class Encoder:
    """src/encoder.rs:7-55"""

    def __init__(self):
        self.range_coder = RangeCoder()
        self.code = deque()

    def peek_code(self):
        return self.code

    def encode(self, pmodel, index):  # :24-37
        out = self.range_coder.param_update(
            pmodel.c_freq(index), pmodel.cum_freq(index), pmodel.total_freq()
        )
        n = len(out)
        self.code.extend(out)
        return n

    def finish(self):  # :40-46
        for _ in range(8):
            self.code.append(self.range_coder.left_shift())
        return self.code


class Decoder:
    """src/decoder.rs:6-55"""

    def __init__(self, code):  # :14-23
        self._range_coder = RangeCoder()
        self._data = 0
        self.buffer = deque(code)
        self.shift_left_buffer(8)

    def range_coder(self):
        return self._range_coder

    def data(self):
        return self._data

    def shift_left_buffer(self, n):  # :31-35
        for _ in range(n):
            self._data = ((self._data << 8) & M64) | self.buffer.popleft()

    def decode(self, pmodel):  # :38-54
        idx = pmodel.find_index(self)
        n = len(
            self._range_coder.param_update(
                pmodel.c_freq(idx), pmodel.cum_freq(idx), pmodel.total_freq()
            )
        )
        self.shift_left_buffer(n)
        return idx


class FreqTable:
    """examples/sample_impl.rs:10-70"""

    def __init__(self, alphabet_count):  # :48-53
        self._total = 0
        self.c = [0] * alphabet_count
        self.cum = [0] * alphabet_count

    @classmethod
    def from_tables(cls, c, cum, total):
        t = cls(len(c))
        t.c = list(c)
        t.cum = list(cum)
        t._total = total
        return t

    def alphabet_count(self):
        return len(self.c)

    def add_alphabet_freq(self, i):  # :58-60
        self.c[i] += 1

    def calc_cum(self):  # :61-69
        cum_total = 0
        for i in range(len(self.c)):
            self.cum[i] = cum_total
            cum_total = (cum_total + self.c[i]) & M32
        self._total = cum_total

    def c_freq(self, i):
        return self.c[i]

    def cum_freq(self, i):
        return self.cum[i]

    def total_freq(self):
        return self._total

    def find_index(self, decoder):  # :27-45
        rc = decoder.range_coder()
        rfreq = (decoder.data() - rc.lower_bound) // rc.range_par_total(self._total)
        left, right = 0, self.alphabet_count() - 1
        while left < right:
            mid = (left + right) // 2
            if self.cum[mid + 1] <= rfreq:
                left = mid + 1
            else:
                right = mid
        return left


def encode(symbols, c, cum, total):
    """Whole `Encoder` run: bytes of `finish()`."""
    t = FreqTable.from_tables(c, cum, total)
    e = Encoder()
    for s in symbols:
        e.encode(t, s)
    return bytes(e.finish())


def decode(code, n, c, cum, total):
    t = FreqTable.from_tables(c, cum, total)
    d = Decoder(code)
    return [d.decode(t) for _ in range(n)]


class AdaptiveFreqTable(FreqTable):
    """f4 (build-defined, see oracle/rc_oracle.h): a FreqTable the caller updates between calls --
    counts start at 1, the coded symbol gains `inc`, all counts are halved (rounding up) when the total
    would pass `limit`; calc_cum() after every change (examples/sample_impl.rs:61-69)."""

    def __init__(self, alphabet_count, inc, limit):
        super().__init__(alphabet_count)
        assert inc >= 1 and alphabet_count <= limit and limit + inc <= M32
        self.inc, self.limit = inc, limit
        self.c = [1] * alphabet_count
        self.calc_cum()

    def update(self, s):
        self.c[s] += self.inc
        if self._total + self.inc > self.limit:
            self.c = [(x + 1) >> 1 for x in self.c]
        self.calc_cum()


def adaptive_encode(symbols, K, inc, limit):
    t = AdaptiveFreqTable(K, inc, limit)
    e = Encoder()
    for s in symbols:
        e.encode(t, s)  # the table as it is before the update
        t.update(s)
    return bytes(e.finish())


def adaptive_decode(code, n, K, inc, limit):
    t = AdaptiveFreqTable(K, inc, limit)
    d = Decoder(code)
    out = []
    for _ in range(n):
        s = d.decode(t)
        t.update(s)
        out.append(s)
    return out


def sample_impl():
    """examples/sample_impl.rs:72-128 -- returns (table, code, decoded)."""
    test_data = [2, 1, 1, 4, 1, 4, 2, 1, 0, 1, 5, 9, 8, 7, 6, 5]
    sd = FreqTable(10)
    for i in test_data:
        sd.add_alphabet_freq(i)
    sd.calc_cum()
    enc = Encoder()
    for i in test_data:
        enc.encode(sd, i)
    code = bytes(enc.finish())
    dec = Decoder(code)
    decoded = [dec.decode(sd) for _ in test_data]
    assert decoded == test_data
    return sd, code, decoded


if __name__ == "__main__":
    sd, code, decoded = sample_impl()
    print("output : 0x" + "".join("%x" % b for b in code))
    print("length : %dbyte" % len(code))
    print(code.hex())
