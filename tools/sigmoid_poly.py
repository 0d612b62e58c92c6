"""Design-time check (numpy, float32) of an FMA-pipe sigmoid for the LSTM epilogue (docs/ROUND2_PLAN.md item 4):
sigma(x) = 1 / (1 + 2^(-x*log2 e)) with 2^y from a Cody-Waite split + degree-3/4 polynomial and the reciprocal from an
integer-trick seed + Newton steps - no MUFU.  Prints the max absolute error against float64 and the instruction count.

    python tools/sigmoid_poly.py
"""
import numpy as np

f32 = np.float32


def exp2_poly(y, degree):
    y = np.clip(y, f32(-126), f32(126)).astype(f32)
    magic = f32(12582912.0)                      # 1.5 * 2^23: (y + magic) - magic rounds to nearest integer
    n = ((y + magic) - magic).astype(f32)
    f = (y - n).astype(f32)                      # [-0.5, 0.5]
    if degree == 3:
        c = [f32(0.99992448), f32(0.69312102), f32(0.24264008), f32(0.05592204)]      # Chebyshev fit on [-0.5, 0.5]: 7.8e-5
    else:
        c = [f32(1.0000001), f32(0.69312102), f32(0.24022107), f32(0.05592204), f32(0.00967604)]       # 2.7e-6
    p = c[-1]
    for k in reversed(c[:-1]):
        p = (p * f + k).astype(f32)
    bits = p.view(np.int32) + (n.astype(np.int32) << 23)          # scale by 2^n through the exponent field
    return bits.view(f32)


def recip_newton(d, steps):
    r = (np.int32(0x7EF311C7) - d.view(np.int32)).view(f32)       # ~12 % relative error seed
    for _ in range(steps):
        r = (r * (f32(2.0) - d * r)).astype(f32)
    return r


def sigmoid_fma(x, degree, steps):
    e = exp2_poly((-x * f32(1.4426950408889634)).astype(f32), degree)
    return recip_newton((f32(1.0) + e).astype(f32), steps)


if __name__ == "__main__":
    x = np.linspace(-20, 20, 2_000_001).astype(f32)
    ref = 1.0 / (1.0 + np.exp(-x.astype(np.float64)))
    for degree in (3, 4):
        for steps in (1, 2, 3):
            err = np.abs(sigmoid_fma(x, degree, steps).astype(np.float64) - ref).max()
            instrs = 2 + 3 + degree + 2 + 1 + 1 + 2 * steps     # clamp, split, Horner, exponent add, 1+e, seed, Newton
            print(f"degree {degree}, {steps} Newton steps: max |err| = {err:.2e}  (~{instrs} FMA/ALU-pipe instructions)")
    print("for scale: tanh.approx.f32 has ~5e-4 relative error; fp16 h rounding is 4.9e-4")
