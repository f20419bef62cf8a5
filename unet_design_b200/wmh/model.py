"""Drop-in for wmh/model.py (the segmentation U-Net with the Haar encoder, a modified copy of pdearena's
`Unetbase_G`; wmh/model.py:1-3): 2 input modalities, 1 mask channel + Sigmoid (:215, :247-254), and the odd
extent handling 200 -> 100 -> 50 -> 25 -> 13 through the zero-extended Haar with the decoder crop `h[:, :, 1:, 1:]`
at the first up level (:146-155; the non-DWT arm replicate-pads instead).  Block classes are shared with
`unet_design_b200.pdearena.modules.twod_unetbase`."""
from __future__ import annotations

from ..pdearena.modules.twod_unetbase import (ConvBlock, DWTBlock, Down_G, FullResnetConvBlock,  # noqa: F401
                                              PartialResnetConvBlock, Up_G, _UnetbaseGCore)
from ..pdearena.modules.activations import resolve


class Unetbase_G(_UnetbaseGCore):
    """wmh `Unetbase_G` (wmh/model.py:165-296): forward(x[B,2,H,W], n_levels_used=None) -> [B,1,H,W] in (0,1)."""

    def __init__(self, hidden_channels: int, activation="gelu", dwt_encoder=False, up_fct="interpolate_nearest",
                 n_extra_resnet_layers=0, multi_res_loss=False, sequ_mode=False, no_skip_connection=False,
                 no_down_up=False, dwt_mode="zero", dwt_wave="haar") -> None:
        super().__init__()
        self.hidden_channels = hidden_channels
        self.activation = resolve(activation)
        self.dwt_encoder, self.up_fct, self.n_extra_resnet_layers = dwt_encoder, up_fct, n_extra_resnet_layers
        self.multi_res_loss, self.sequ_mode, self.no_skip_connection = multi_res_loss, sequ_mode, no_skip_connection
        self.no_down_up, self.dwt_mode, self.dwt_wave = no_down_up, dwt_mode, dwt_wave
        self._build(2, hidden_channels, 1, activation, final_sigmoid=True, crop_finest=True)

    def forward(self, x, n_levels_used=None):
        if n_levels_used is None:
            n_levels_used = self.n_levels
        return self._run(x, n_levels_used)
