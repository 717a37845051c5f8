"""The algorithmic FLOP count used by bench.py's FP64 roofline is an instrumented count of the
oracle (reference arithmetic), not a hand estimate."""
import re


def test_instrumented_flop_count_matches_bench_constant():
    from oracle.count_flops import count

    c = count(n=2000)
    flop = c["flop_add_mul_div_sqrt"]
    assert 3000 < flop < 5000 and c["div"] > 200 and c["sqrt"] > 80
    src = open(__import__("os").path.join(__import__("os").path.dirname(__file__), "..", "bench.py")).read()
    const = float(re.search(r"^ALG_FLOP_TRACE = ([0-9.]+)", src, re.M).group(1))
    assert abs(const - flop) / flop < 0.02, (const, flop)
