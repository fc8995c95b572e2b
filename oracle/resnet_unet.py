"""Oracle restatement of the reference's in-tree ResNet encoder/decoder (TEST INFRASTRUCTURE).

Follows SU/UArchModel/resnet_unet.py:36-44 (convrelu), :134-213 (__init__), :215-300 (forward).
Pinned by tests/golden/resnet_unet_reference.npz, produced by running the reference's own file
(oracle/make_golden.py): with the same seed this module builds bit-identical weights (same
construction order) and must reproduce the golden logits exactly.
The torchvision backbone is random-init (`weights=None`): no network here, and BASELINE config
3 says random-init encoder.
"""
import torch
import torch.nn as nn
import torchvision


def convrelu(cin, cout, kernel, padding):  # resnet_unet.py:36-44
    return nn.Sequential(nn.Conv2d(cin, cout, kernel, padding=padding), nn.ReLU(inplace=True))


class ResNetUNet(nn.Module):
    def __init__(self, n_class, resnet_model):
        super().__init__()
        self.resnet_model = resnet_model
        if resnet_model == 18:
            self.base_model = torchvision.models.resnet18(weights=None)
        elif resnet_model == 34:
            self.base_model = torchvision.models.resnet34(weights=None)
        else:
            raise ValueError("Only ResNet-18 and ResNet-34 are supported")
        self.base_layers = list(self.base_model.children())
        self.layer0 = nn.Sequential(*self.base_layers[:3])
        self.layer0_1x1 = convrelu(64, 64, 1, 0)
        self.layer1 = nn.Sequential(*self.base_layers[3:5])
        self.layer1_1x1 = convrelu(64, 64, 1, 0)
        self.layer2 = self.base_layers[5]
        self.layer2_1x1 = convrelu(128, 128, 1, 0)
        self.layer3 = self.base_layers[6]
        self.layer3_1x1 = convrelu(256, 256, 1, 0)
        self.layer4 = self.base_layers[7]
        self.layer4_1x1 = convrelu(512, 512, 1, 0)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv_up3 = convrelu(256 + 512, 512, 3, 1)
        self.conv_up2 = convrelu(128 + 512, 256, 3, 1)
        self.conv_up1 = convrelu(64 + 256, 256, 3, 1)
        self.conv_up0 = convrelu(64 + 256, 128, 3, 1)
        self.conv_original_size0 = convrelu(3, 64, 3, 1)
        self.conv_original_size1 = convrelu(64, 64, 3, 1)
        self.conv_original_size2 = convrelu(64 + 128, 64, 3, 1)
        self.conv_last = nn.Conv2d(64, n_class, 1)

    def forward(self, input):  # resnet_unet.py:252-298
        x_original = self.conv_original_size0(input)
        x_original = self.conv_original_size1(x_original)
        layer0 = self.layer0(input)
        layer1 = self.layer1(layer0)
        layer2 = self.layer2(layer1)
        layer3 = self.layer3(layer2)
        layer4 = self.layer4(layer3)
        layer4 = self.layer4_1x1(layer4)
        x = self.upsample(layer4)
        layer3 = self.layer3_1x1(layer3)
        x = torch.cat([x, layer3], dim=1)
        x = self.conv_up3(x)
        x = self.upsample(x)
        layer2 = self.layer2_1x1(layer2)
        x = torch.cat([x, layer2], dim=1)
        x = self.conv_up2(x)
        x = self.upsample(x)
        layer1 = self.layer1_1x1(layer1)
        x = torch.cat([x, layer1], dim=1)
        x = self.conv_up1(x)
        x = self.upsample(x)
        layer0 = self.layer0_1x1(layer0)
        x = torch.cat([x, layer0], dim=1)
        x = self.conv_up0(x)
        x = self.upsample(x)
        x = torch.cat([x, x_original], dim=1)
        x = self.conv_original_size2(x)
        return self.conv_last(x)


def normalize(batch, mean, std):
    """SU/utils.py:480-519: per-image (x - mean) / std."""
    mean = mean.view(-1, 1, 1)
    std = std.view(-1, 1, 1)
    return torch.cat([((batch[i] - mean) / std).unsqueeze(0) for i in range(len(batch))], 0)
