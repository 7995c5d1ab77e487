"""TransD: dynamic projection e' = normalize(e + (e.e_p) r_p) before the TransE energy
(reference openke/module/model/TransD.py:8-129).  Only dim_e == dim_r is supported: it is the only
shape the reference's experiments use, and the reference's own zero-padding branch for
dim_e < dim_r (TransD.py:62-76) raises inside torch (wrong keyword to F.pad)."""
from ... import _native as N
from .Model import Model


class TransD(Model):
    _pk_model = N.PK_TRANSD
    _ent_tables = ("ent_embeddings", "ent_transfer")
    _rel_tables = ("rel_embeddings", "rel_transfer")

    def __init__(self, ent_tot, rel_tot, dim_e=100, dim_r=100, p_norm=1, norm_flag=True, margin=None, epsilon=None):
        super().__init__(ent_tot, rel_tot)
        if dim_e != dim_r:
            raise NotImplementedError("TransD on the B200 path needs dim_e == dim_r (got %d, %d)" % (dim_e, dim_r))
        self.dim_e, self.dim_r, self.margin, self.epsilon = dim_e, dim_r, margin, epsilon
        self.norm_flag, self.p_norm = norm_flag, p_norm
        rng = None if margin is None or epsilon is None else (margin + epsilon) / dim_e
        # creation order = reference TransD.py:18-21 (it fixes the torch RNG stream of the init)
        self._init_tables(self.table_specs(ent_tot, rel_tot, dim_e=dim_e, dim_r=dim_r), margin, epsilon,
                          {"ent_embedding_range": rng, "rel_embedding_range": rng})

    @classmethod
    def table_specs(cls, ent_tot, rel_tot, dim_e=100, dim_r=100, **_):
        return [("ent_embeddings", ent_tot, dim_e), ("rel_embeddings", rel_tot, dim_r),
                ("ent_transfer", ent_tot, dim_e), ("rel_transfer", rel_tot, dim_r)]
