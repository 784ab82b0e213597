"""PairList — the argument type of ``BaseModel.forward`` (lib/dataset/list_pair.py:3-57).

Same interface as the reference (``features``, ``extra_fields``, ``add_field/get_field/
has_field/fields``, ``to``, ``__getitem__``, ``__len__``, ``copy_with_fields``).  The
reference fields are ``tracklet_pairs``, ``track_cls_logits``, ``num_tracklets``, ``ious``,
``track_ids`` (vrdataset.py:61-83); [SPEC] s1 adds the dense tracklet inputs from which the
pair stage builds everything else: ``boxes [N,T,4]``, ``span [N,2]``, ``motion [N,4000]``.
``features`` may be ``None`` when the rows are to be constructed on the GPU.
"""
from __future__ import annotations

import torch


class PairList:
    def __init__(self, feat):
        self.features = feat
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, feat):
        for k, v in feat.extra_fields.items():
            self.extra_fields[k] = v

    def to(self, device):
        feat = PairList(self.features.to(device) if self.features is not None else None)
        for k, v in self.extra_fields.items():
            if hasattr(v, "to"):
                v = v.to(device)
            feat.add_field(k, v)
        return feat

    def __getitem__(self, item):
        feat = PairList(self.features[item])
        for k, v in self.extra_fields.items():
            feat.add_field(k, v[item])
        return feat

    def __len__(self):
        if self.features is not None:
            return self.features.shape[0]
        n = int(self.extra_fields["num_tracklets"])
        return n * (n - 1)

    def copy_with_fields(self, fields, skip_missing=False):
        feat = PairList(self.features)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for field in fields:
            if self.has_field(field):
                feat.add_field(field, self.get_field(field))
            elif not skip_missing:
                raise KeyError("Field '{}' not found in {}".format(field, self))
        return feat

    def __repr__(self):
        return self.__class__.__name__ + "(num_feats={})".format(len(self))

    # [SPEC] s1 ---------------------------------------------------------------------------
    @classmethod
    def from_tracklets(cls, boxes, span, track_cls_logits, motion=None, features=None):
        """PairList carrying the dense tracklet inputs (no precomputed feature rows)."""
        boxes = torch.as_tensor(boxes)
        pl = cls(features)
        n = int(boxes.shape[0])
        pl.add_field("boxes", boxes)
        pl.add_field("span", torch.as_tensor(span))
        pl.add_field("track_cls_logits", torch.as_tensor(track_cls_logits))
        if motion is not None:
            pl.add_field("motion", torch.as_tensor(motion))
        pl.add_field("num_tracklets", n)
        return pl

    def has_tracklets(self):
        return self.has_field("boxes") and self.has_field("span")
