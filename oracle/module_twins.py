"""TEST / BASELINE INFRASTRUCTURE ONLY -- "oracle twins" of the drop-in modules: the product's parameter containers (same
``state_dict`` keys, never executed) whose ``forward`` is the oracle's functional restatement of the reference module
(fp32 eager: dense epipolar bias per call, mask replicated over the heads, plain softmax attention).  A UNet built with
these factories is the reference's own PyTorch path on the same parameter tree: what the product is compared with in
tests/test_gpu_sampler.py / tests/test_gpu_speedup.py and timed beside in the ``reference_gpu`` leg of bench.py.
Nothing under cvd_b200/ imports this module."""


def make():
    """Returns (OracleEpi, OracleMotion) classes (built lazily: importing the product package must not import oracle/)."""
    import cvd_b200
    from . import epi_oracle as O

    class OracleEpi(cvd_b200.EpiModule):
        def forward(self, hidden_states, F_mats=None, H_mats=None, temb=None, **kw):
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            return O.epi_module(hidden_states, sd, 8, F_mats, 256)

    class OracleMotion(cvd_b200.VanillaTemporalModule):
        def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                    cross_attention_kwargs={}):
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            p = "temporal_transformer.transformer_blocks.0.attention_blocks.0.processor.qkv_merge."
            merges = [(sd[p + "weight"], sd[p + "bias"]), None] if p + "weight" in sd else None
            return O.temporal_module(hidden_states, sd, 8, cross_attention_kwargs.get("pose_feature"), merges)

    return OracleEpi, OracleMotion


def sync_unet_factories():
    """Factories with the signature ``cvd_b200.sync_unet.SyncPathUNet`` expects: (channels, kwargs)."""
    import cvd_b200
    OracleEpi, OracleMotion = make()

    def epi_factory(c, kw):
        return OracleEpi(in_channels=c, **kw)

    def motion_factory(c, kw):
        m = OracleMotion(in_channels=c, **kw)
        blk = m.temporal_transformer.transformer_blocks[0]
        blk.attention_blocks[0].set_processor(cvd_b200.PoseAdaptorAttnProcessor(
            hidden_size=c, pose_feature_dim=c, query_condition=True, key_value_condition=True, scale=1.0))
        blk.attention_blocks[1].set_processor(cvd_b200.AttnProcessor(attn_type="temporal"))
        return m

    return dict(epi_factory=epi_factory, motion_factory=motion_factory)


def unet_factories():
    """Factories with the reference's own factory signatures (``get_motion_module(in_channels, type, kwargs)``,
    ``get_epi_module(in_channels, kwargs)``) for ``cvd_b200.unet.UNet3DConditionModelPoseCond``; the pose adaptor is
    installed afterwards by ``set_all_attn_processor`` exactly as on the product UNet."""
    OracleEpi, OracleMotion = make()
    return dict(epi_factory=lambda c, kw: OracleEpi(in_channels=c, **kw),
                motion_factory=lambda c, kind, kw: OracleMotion(in_channels=c, **kw))
