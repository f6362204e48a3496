"""TEST INFRASTRUCTURE: an HF ``BlipForConditionalGeneration`` behind the ``caption_loss`` protocol of ``BlipEngine``, so logic
tests can put the library module where the product wants its native executor (the product itself refuses HF modules)."""
import torch


class HFBlipComparator:
    def __init__(self, model):
        self.model = model
        for p in model.parameters():
            p.requires_grad = False

    def caption_loss(self, pixel_values, input_ids, attention_mask, labels):
        dt = next(self.model.parameters()).dtype
        with torch.autocast(pixel_values.device.type, dtype=dt, enabled=dt != torch.float32):
            return self.model(pixel_values=pixel_values.to(dt), input_ids=input_ids, attention_mask=attention_mask, labels=labels).loss
