"""`pytorch3d.structures.Meshes` -- the subset GauSTAR touches (gaustar_scene/sugar_model.py:570-600, gaustar_trainers/
refine.py:291), restated from the published behaviour of pytorch3d 0.7.4 (PARITY UNPINNED: the package is not in this
image).  Batches of meshes are supported as lists; "packed" tensors concatenate the meshes with vertex indices offset.

Conventions kept: `edges_packed()` lists every undirected edge once as (min, max), sorted lexicographically;
`faces_packed_to_edges_packed()[f] = (edge opposite vertex 0, opposite vertex 1, opposite vertex 2)` of face f; face
normals follow the winding (v1 - v0) x (v2 - v0); areas are half the norm of that cross product.
"""
import torch


class Meshes:
    def __init__(self, verts=None, faces=None, textures=None):
        if isinstance(verts, torch.Tensor):
            verts = [verts] if verts.dim() == 2 else list(verts)
        if isinstance(faces, torch.Tensor):
            faces = [faces] if faces.dim() == 2 else list(faces)
        if len(verts) != len(faces):
            raise ValueError("verts and faces must describe the same number of meshes")
        self._verts_list = [v for v in verts]
        self._faces_list = [f.to(torch.int64) for f in faces]
        self.textures = textures
        self.device = self._verts_list[0].device if self._verts_list else torch.device("cpu")
        self._cache = {}

    def __len__(self):
        return len(self._verts_list)

    def isempty(self):
        return len(self) == 0 or all(v.numel() == 0 for v in self._verts_list)

    def verts_list(self):
        return self._verts_list

    def faces_list(self):
        return self._faces_list

    def num_verts_per_mesh(self):
        return torch.tensor([v.shape[0] for v in self._verts_list], dtype=torch.int64, device=self.device)

    def num_faces_per_mesh(self):
        return torch.tensor([f.shape[0] for f in self._faces_list], dtype=torch.int64, device=self.device)

    def verts_packed(self):
        return torch.cat(self._verts_list, 0) if len(self) else torch.zeros(0, 3)

    def mesh_to_verts_packed_first_idx(self):
        n = self.num_verts_per_mesh()
        return torch.cumsum(n, 0) - n

    def faces_packed(self):
        if "faces" not in self._cache:
            off = self.mesh_to_verts_packed_first_idx()
            self._cache["faces"] = torch.cat([f + o for f, o in zip(self._faces_list, off)], 0) if len(self) else torch.zeros(0, 3, dtype=torch.int64)
        return self._cache["faces"]

    def verts_packed_to_mesh_idx(self):
        return torch.repeat_interleave(torch.arange(len(self), device=self.device), self.num_verts_per_mesh())

    def faces_packed_to_mesh_idx(self):
        return torch.repeat_interleave(torch.arange(len(self), device=self.device), self.num_faces_per_mesh())

    def _edges(self):
        if "edges" not in self._cache:
            F = self.faces_packed()
            V = max(int(self.verts_packed().shape[0]), 1)
            e = torch.cat([F[:, [1, 2]], F[:, [2, 0]], F[:, [0, 1]]], 0)  # opposite vertex 0, 1, 2
            e = torch.stack([e.min(1).values, e.max(1).values], 1)
            u, inv = torch.unique(e[:, 0] * V + e[:, 1], sorted=True, return_inverse=True)
            self._cache["edges"] = torch.stack([u // V, u % V], 1)
            self._cache["f2e"] = inv.reshape(3, F.shape[0]).t().contiguous()
        return self._cache["edges"], self._cache["f2e"]

    def edges_packed(self):
        return self._edges()[0]

    def faces_packed_to_edges_packed(self):
        return self._edges()[1]

    def _face_cross(self):
        v, f = self.verts_packed(), self.faces_packed()
        v0, v1, v2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        return torch.cross(v1 - v0, v2 - v0, dim=1)

    def faces_areas_packed(self):
        return 0.5 * self._face_cross().norm(dim=1)

    def faces_normals_packed(self):
        c = self._face_cross()
        return c / c.norm(dim=1, keepdim=True).clamp(min=1e-6)

    def faces_normals_list(self):
        return list(torch.split(self.faces_normals_packed(), [int(n) for n in self.num_faces_per_mesh()], 0))

    def verts_normals_packed(self):
        """Area-weighted average of the incident face normals, normalised."""
        v, f = self.verts_packed(), self.faces_packed()
        c = self._face_cross()
        n = torch.zeros_like(v)
        for k in range(3):
            n = n.index_add(0, f[:, k], c)
        return torch.nn.functional.normalize(n, eps=1e-6, dim=1)
