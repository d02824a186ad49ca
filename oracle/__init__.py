"""CPU oracle for the hot path of ark-crypto-primitives.  TEST INFRASTRUCTURE ONLY.

Nothing in the product (``crypto_primitives_b200/``, ``include/``) may import,
link or execute anything under ``oracle/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may, and there only as the checker / the timed CPU baseline.

Two restatements live here:

* ``oracle.poseidon`` / ``oracle.jubjub`` / ``oracle.pedersen`` / ``oracle.merkle``
  -- Python big-integer restatement of the reference algorithms (slow, exact).
* ``oracle/cref/oracle_ref.c`` -- plain-C restatement (4x64-bit Montgomery,
  pthread data-parallel like the reference's rayon path), built by
  ``oracle/Makefile`` into ``oracle/_ref/liboracle_ref.so``; bit-exact with the
  Python one and used at sizes Python cannot finish, and as the timed CPU
  baseline.

Pinning status:
  Poseidon (Grain LFSR, default params, sponge, CRH, two-to-one): PINNED by the
  reference's own known-answer tests (grain_lfsr.rs:190-218, traits.rs:163-358,
  sponge/poseidon/mod.rs:381-404) -- transcribed in tests/golden/.
  Pedersen CRH / commitment and Merkle roots: PARITY UNPINNED -- the reference
  holds no constants for them (only native==gadget and prove/verify round trips
  under an RNG that cannot be replayed here; no Rust toolchain in this image).
"""
