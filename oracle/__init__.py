"""CPU oracle for the IDEAS GAN-conv hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain, slow, CPU restatement of what the
reference (Lemok00/IDEAS, ``/root/reference``) computes on the hot path.  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``ideas_b200``) never imports this package and has no CPU fallback.

Parity status: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` unmodified) and committed under ``tests/golden/``; the
RNG-free bit-path known answers of SURVEY.md App. E are pinned as literals in
``tests/test_oracle_bits.py``.

Modules
  functional.py  op-level restatements (bias+LeakyReLU, upfirdn2d, equalised and
                 modulated convolutions), torch-CPU fp32, differentiable.
  nets.py        the seven IDEAS networks as pure functions of a state_dict,
                 plus the parameter spec / seeded initialiser.
  bits.py        message <-> secret-tensor mapping and BER in numpy.
  train_step.py  one IDEAS training iteration (train.py:33-221) over the
                 functional nets, used as the timed CPU baseline.
"""
