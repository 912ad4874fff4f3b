"""CPU oracle for the NeuralCodecs codec hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (PyTorch fp32 / fp64, same ATen op family that TorchSharp
binds) restatement of the reference's C# modules for DAC / SNAC / Encodec
encode+decode.  It exists to *check* the CUDA engine and to serve as the timed
"reference CPU path" baseline in ``bench.py``.  Nothing under
``neuralcodecs_b200/`` may import it: the product path is the C-ABI library and
fails loudly if that library is missing.

PARITY UNPINNED: the reference repository ships no tests, fixtures or golden
vectors for this path and cannot be executed here
(C#/.NET + TorchSharp, no dotnet toolchain).  The oracle is pinned instead
against (i) committed fixtures produced by ``transformers`` DAC/Encodec models
(structure: paddings, strides, key mapping; see tests/golden/make_golden.py) and
(ii) hand-written numpy loop implementations of the individual ops on small
shapes (tests/test_oracle_*.py).
"""
