"""TEST INFRASTRUCTURE: a CPU stage-op provider for alphafold2_b200.parallel.sharded_evoformer_forward built on the
oracle math (fp64).  It lets the world_size-2 gloo test check the sharding SCHEDULE (slicing, all-to-all layouts,
operand gathers, collective order) without a GPU.  Never used by the product."""
import torch
import torch.nn.functional as F

from oracle import evoformer_oracle as O


def _w(mod, dt=torch.float64):
    return {k: v.detach().to(dt) for k, v in mod.state_dict().items()}


class OracleStageOps:
    def pair_bias(self, ax, x_rows):
        w = ax.edges_to_attn_bias[0].weight.detach().to(x_rows.dtype)
        return torch.einsum("rjd,hd->hrj", x_rows, w).contiguous()

    def axial_attention_(self, ax, x, bias, mask, row_attn):
        w = _w(ax, x.dtype)
        xn = F.layer_norm(x, (x.shape[-1],), w["norm.weight"], w["norm.bias"], 1e-5)
        if row_attn:
            xf, mf = xn, mask
        else:
            xf, mf = xn.transpose(0, 1), (None if mask is None else mask.transpose(0, 1))
        out = O.attention(w, "attn.", xf.contiguous(), ax.attn.heads, mf, None if bias is None else bias[None])
        x += out if row_attn else out.transpose(0, 1)

    def feed_forward_(self, ff, x):
        x += O.feed_forward(_w(ff, x.dtype), "", x)

    def outer_project(self, om, m_cols, mask_cols):
        w = _w(om, m_cols.dtype)
        mn = F.layer_norm(m_cols, (m_cols.shape[-1],), w["norm.weight"], w["norm.bias"], 1e-5)
        L = F.linear(mn, w["left_proj.weight"], w["left_proj.bias"])
        R = F.linear(mn, w["right_proj.weight"], w["right_proj.bias"])
        if mask_cols is not None:
            mk = mask_cols[..., None].to(mn.dtype)
            L, R = L * mk, R * mk
        return torch.cat([L.permute(2, 0, 1), R.permute(2, 0, 1)], 0).contiguous()     # [2d, S, nl]

    def outer_contract_(self, om, x_rows, L, Rg, msa_mask_full, row0, pieces):
        w = _w(om, x_rows.dtype)
        d, S, rows = L.shape
        R = Rg.view(pieces, d, S, -1).permute(1, 2, 0, 3).reshape(d, S, -1)              # [d, S, N]
        outer = torch.einsum("csi,csj->ijc", L, R) / S
        if msa_mask_full is not None:
            mf = msa_mask_full.to(torch.float32)
            cnt = torch.einsum("si,sj->ij", mf[:, row0:row0 + rows], mf)
            outer = outer / (cnt + om.eps)[..., None].to(outer.dtype)
        x_rows += F.linear(outer, w["proj_out.weight"], w["proj_out.bias"])

    def tri_project(self, tm, x_loc, mask_loc):
        w = _w(tm, x_loc.dtype)
        xn = F.layer_norm(x_loc, (x_loc.shape[-1],), w["norm.weight"], w["norm.bias"], 1e-5)
        lin = lambda k: F.linear(xn, w[k + ".weight"], w[k + ".bias"])  # noqa: E731
        L, R = lin("left_proj"), lin("right_proj")
        if mask_loc is not None:
            mk = mask_loc[..., None].to(xn.dtype)
            L, R = L * mk, R * mk
        L = L * lin("left_gate").sigmoid()
        R = R * lin("right_gate").sigmoid()
        G = lin("out_gate").sigmoid().reshape(-1, x_loc.shape[-1])
        return L.permute(2, 0, 1).contiguous(), R.permute(2, 0, 1).contiguous(), G

    def tri_contract_(self, tm, x_loc, L, Rg, gate, ingoing, pieces):
        w = _w(tm, x_loc.dtype)
        d = L.shape[0]
        if not ingoing:
            R = Rg.view(pieces, d, Rg.shape[1], Rg.shape[2]).permute(1, 0, 2, 3).reshape(d, -1, Rg.shape[2])   # [d, N(j), K]
            out = torch.einsum("cik,cjk->ijc", L, R)
        else:
            R = Rg.view(pieces, d, Rg.shape[1], Rg.shape[2]).permute(1, 2, 0, 3).reshape(d, Rg.shape[1], -1)   # [d, K, N(i)]
            out = torch.einsum("cki,ckj->ijc", R, L)
        out = F.layer_norm(out, (d,), w["to_out_norm.weight"], w["to_out_norm.bias"], 1e-5) * gate.view(out.shape)
        x_loc += F.linear(out, w["to_out.weight"], w["to_out.bias"])
