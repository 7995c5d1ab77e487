"""TransH: entities are projected onto the relation's hyperplane, e - (e.w^)w^, before the TransE
energy (reference openke/module/model/TransH.py:8-92)."""
from ... import _native as N
from .Model import Model


class TransH(Model):
    _pk_model = N.PK_TRANSH
    _rel_tables = ("rel_embeddings", "norm_vector")

    def __init__(self, ent_tot, rel_tot, dim=100, p_norm=1, norm_flag=True, margin=None, epsilon=None):
        super().__init__(ent_tot, rel_tot)
        self.dim, self.margin, self.epsilon = dim, margin, epsilon
        self.norm_flag, self.p_norm = norm_flag, p_norm
        rng = None if margin is None or epsilon is None else (margin + epsilon) / dim
        self._init_tables(self.table_specs(ent_tot, rel_tot, dim=dim), margin, epsilon, {"embedding_range": rng})

    @classmethod
    def table_specs(cls, ent_tot, rel_tot, dim=100, **_):
        return [("ent_embeddings", ent_tot, dim), ("rel_embeddings", rel_tot, dim), ("norm_vector", rel_tot, dim)]
