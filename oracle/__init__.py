"""CPU oracle for the whisper.apr mel + encoder hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference's Rust path (file:line cited on each function).  It may be
imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- always as the
checker, never as the thing measured or shipped.  The product path
(``whisper_apr_b200``) must never import it.

Pinning status (see DESIGN.md "Oracle"):
  * mel: pinned against the reference's in-tree golden vectors
    (tests/golden/ref_a_audio.bin -> ref_c_mel_numpy.bin, bit-exact with the
    symmetric window that produced the golden; 0.0138 max-abs with the
    reference's own periodic window, which is the reference's own documented
    tolerance band) and its frame-count known-answer tests.
  * .apr format / int8 / int4: pinned against the reference's unit-test KATs
    (header bytes, descriptor layout, CRC-32 check values, quantiser bounds).
  * encoder values: the reference ships NO numeric golden for encoder states
    and cannot be compiled here (no Rust toolchain) -> "parity unpinned" for
    absolute encoder values; the restatement is pinned structurally (conv
    length KATs, flash == naive softmax attention, LayerNorm/GELU identities).
"""
