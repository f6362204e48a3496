"""CPU/GPU-eager *oracle* for the CoMat per-step training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``comat_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and there only as the checker / the timed CPU baseline.

What it is: a plain-PyTorch restatement of the third-party arithmetic the reference calls
(diffusers 0.22-0.25 ``UNet2DConditionModel`` / ``AutoencoderKL.decode`` / ``DDPMScheduler`` /
``LoRALinearLayer``; none of that source is vendored under /root/reference) plus restatements of
the reference's own loss code (``attn_utils/tc_loss_utils.py``, ``attn_utils/tc_attn_utils.py``,
``attr_concen_utils/gsam_interface.py:get_mask_loss``, ``concept_mat_utils/caption_blip.py:score``,
``training_utils/gan_sdxl.py``, the step assembly of ``training_script.py:556-664``).

Pinning status (see DESIGN.md "Oracle"):
  * the restated CoMat-owned functions are pinned against the reference's OWN modules imported in the
    build container (``oracle/pin_against_reference.py`` -> ``tests/golden/*.pt``);
  * the restated third-party layer is **parity unpinned** by upstream tests (the reference has no
    tests or golden vectors); it is anchored structurally: the four published
    parameter totals, hookability by the reference's own ``register_attention_control`` and the
    captured-map key set/lengths.
"""
