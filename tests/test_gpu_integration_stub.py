"""The ctypes stub printed in INTEGRATION.md section 2 is run as written (only the library path is made absolute) and compared
with the oracle's SpiralConv (models.py:34-53): the documented binding is a tested binding."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import spiral_oracle as so
from tests.helpers import TOL_BF16, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_matches_oracle():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.findall(r"```python\n(.*?)```", text, re.S)[-1]
    assert "def spiralconv_forward" in code
    code = code.replace('"libshb200.so"', repr(os.path.join(ROOT, "semantichuman_b200", "libshb200.so")))
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    g = torch.Generator().manual_seed(0)
    B, V, S, cin, cout = 5, 40, 9, 16, 32
    table = torch.randint(0, V + 1, (V + 1, S), generator=g).numpy()
    table[-1] = V
    x = torch.randn(B, V + 1, cin, generator=g)
    x[:, -1] = 0
    w = torch.randn(cout, S * cin, generator=g) / (S * cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    y = ns["spiralconv_forward"](x.cuda(), table, w.cuda(), b.cuda(), "elu")
    torch.cuda.synchronize()
    ref = so.spiral_conv(x.bfloat16().float(), torch.from_numpy(table).long()[None].repeat(B, 1, 1), w.bfloat16().float(), b, "elu")
    assert relerr(y.cpu(), ref) <= TOL_BF16
