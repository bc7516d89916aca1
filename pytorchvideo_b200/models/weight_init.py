"""Parameter initialisation (reference models/weight_init.py:10-92).  Init only - never on the
forward path; parity tests always load explicit weights."""
import torch.nn as nn


def _kaiming_fan_out(conv):
    nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
    if conv.bias is not None:
        nn.init.zeros_(conv.bias)


def init_net_weights(model, init_std=0.01, style="resnet"):
    assert style in ("resnet", "vit")
    for m in model.modules():
        if style == "resnet":
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                _kaiming_fan_out(m)
            elif isinstance(m, nn.modules.batchnorm._NormBase):
                if m.weight is not None:
                    # the last BN of each residual branch starts at zero (identity blocks)
                    m.weight.data.fill_(0.0 if getattr(m, "block_final_bn", False) else 1.0)
                if m.bias is not None:
                    m.bias.data.zero_()
            if isinstance(m, nn.Linear):
                if getattr(m, "xavier_init", False):
                    nn.init.kaiming_uniform_(m.weight, a=1)
                else:
                    m.weight.data.normal_(mean=0.0, std=init_std)
                if m.bias is not None:
                    m.bias.data.zero_()
        else:
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=init_std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif type(m).__name__ == "SpatioTemporalClsPositionalEncoding":
                for w in m.parameters():
                    nn.init.trunc_normal_(w, std=init_std)
    return model
