"""OneStageModel (mirrors /root/reference/simvg/models/det_seg/one_stage.py:6-26).  Attribute names `vis_enc` / `head`
are part of the contract: tools/train.py:80-88 selects learning-rate groups by the substring `vis_enc`."""
from simvg_b200.models.builder import MODELS, build_fusion, build_head, build_lan_enc, build_vis_enc

from .base import BaseModel


@MODELS.register_module()
class OneStageModel(BaseModel):
    def __init__(self, word_emb, num_token, vis_enc, lan_enc, head, fusion):
        super().__init__()
        self.vis_enc = build_vis_enc(vis_enc)
        if lan_enc is not None:
            self.lan_enc = build_lan_enc(lan_enc, {"word_emb": word_emb, "num_token": num_token})
        if head is not None:
            self.head = build_head(head)
        if fusion is not None:
            self.fusion = build_fusion(fusion)
