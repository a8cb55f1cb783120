"""CPU oracle for the CDSegNet single-step forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cdsegnet_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the
reported CPU baseline -- never as the product path.

Every function cites the reference file:line (relative to the upstream
``QWTforGithub/CDSegNet`` tree) whose behaviour it restates.

Pinning status (see DESIGN.md "Oracle"):
  * space-filling-curve codes, argsort/inverse, patch padding maps, the whole
    PTv3 dual-network wiring: PINNED against outputs of the reference's own
    python sources executed in the authoring container
    (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz|*.pt``).
  * the DefaultSegmentorV2 wrapper (diffusion schedule, q / DDIM samplers, criteria, inference / inference_ddim / forward) --
    ``wrapper_oracle.py``: PINNED by ``tests/golden/make_golden_wrapper.py`` (reference default.py + losses/*.py executed);
  * the training pass (train-mode forward + autograd gradients) -- ``wrapper_oracle.training_loss`` + ``ptv3_oracle.BN_TRAIN``:
    PINNED by ``tests/golden/make_golden_train.py`` (reference forward + loss.backward() executed);
  * test-time GridSample / FNV + ravel hashes / collate / voting -- ``fragments_np.py``: PINNED by
    ``tests/golden/make_golden_fragments.py`` (reference transform.py + datasets/utils.py executed), except the tie order inside a
    voxel, which the reference's unstable ``np.argsort`` leaves unspecified;
  * ``pointops.knn_query`` -- ``knn_np.py``: PINNED on the GPU box against the reference's OWN CUDA kernel, compiled by
    ``oracle/Makefile`` into ``oracle/_ref/libref_knn.so`` from the sources where they lie under /root/reference;
  * spconv ``SubMConv3d`` tap order / weight layout, ``torch_scatter.segment_csr``
    and ``flash_attn`` numerics: third-party packages that are NOT vendored in the
    reference tree -> "parity unpinned" for those three (semantics restated from
    their published behaviour; see oracle/ptv3_oracle.py headers).
"""
