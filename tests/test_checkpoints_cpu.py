"""Flax msgpack import (row N4): container round trip, dtype handling, spec selection.  The format is restated from
flax.serialization's published definition - no real checkpoint exists in this image, so this pins self-consistency and
the byte layout of the ndarray extension type only."""
import msgpack
import numpy as np
import pytest

from latent_diffusion_planning_b200 import checkpoints as CK, params as P


def test_ndarray_extension_byte_layout():
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    raw = CK.save_flax_msgpack({"w": a})
    top = msgpack.unpackb(raw, raw=False)
    ext = top["w"]
    assert isinstance(ext, msgpack.ExtType) and ext.code == 1
    shape, dtype, buf = msgpack.unpackb(ext.data, raw=False)
    assert shape == [2, 3] and dtype == "float32" and buf == a.tobytes()


def test_roundtrip_nested_tree_and_dtypes(tmp_path):
    g = np.random.default_rng(0)
    tree = {"params": {"encoder": {"conv_in": {"kernel": g.normal(size=(3, 3, 3, 8)).astype(np.float32), "bias": g.normal(size=8).astype(np.float16)}},
                       "step": np.asarray(7, np.int32)}}
    CK.save_flax_msgpack(tree, tmp_path / "m.msgpack")
    back = CK.load_flax_msgpack(tmp_path / "m.msgpack")
    assert np.array_equal(back["params"]["encoder"]["conv_in"]["kernel"], tree["params"]["encoder"]["conv_in"]["kernel"])
    assert back["params"]["encoder"]["conv_in"]["bias"].dtype == np.float16 and int(back["params"]["step"]) == 7
    # bfloat16 leaves (what a bf16-trained Flax model stores) widen to float32 through the bit pattern
    x = np.array([1.0, -2.5, 3.140625], np.float32)
    bits = (x.view(np.uint32) >> 16).astype(np.uint16)
    ext = msgpack.ExtType(1, msgpack.packb(([3], "bfloat16", bits.tobytes()), use_bin_type=True))
    assert np.array_equal(CK.load_flax_msgpack(msgpack.packb({"b": ext}, use_bin_type=True))["b"], x)
    # chunked arrays
    chunked = {"__msgpack_chunked_array__": True, "shape": [2, 3], "chunks": {"0": np.arange(4, dtype=np.float32), "1": np.arange(4, 6, dtype=np.float32)}}
    raw = msgpack.packb({"big": CK._encode(chunked) | {"__msgpack_chunked_array__": True, "shape": [2, 3]}}, use_bin_type=True)
    assert np.array_equal(CK.load_flax_msgpack(raw)["big"], np.arange(6, dtype=np.float32).reshape(2, 3))
    # the layout flax.serialization._chunk really writes: the shape tuple goes through _tuple_to_dict -> {'0': 2, '1': 3}
    flax_chunked = {"__msgpack_chunked_array__": True, "shape": {"0": 2, "1": 3},
                    "chunks": CK._encode({"0": np.arange(4, dtype=np.float32), "1": np.arange(4, 6, dtype=np.float32)})}
    raw = msgpack.packb({"big": flax_chunked}, use_bin_type=True)
    assert np.array_equal(CK.load_flax_msgpack(raw)["big"], np.arange(6, dtype=np.float32).reshape(2, 3))


def test_vae_file_splits_into_encoder_and_decoder_specs(tmp_path):
    blocks = (32, 64)
    enc = P.init_params(P.vae_encoder_spec(blocks, 3, 4, 1), seed=1)
    dec = P.init_params(P.vae_decoder_spec(blocks, 3, 4, 1), seed=2)
    tree = {"params": P.nest({**enc, **dec})}
    (tmp_path / "vae").mkdir()
    CK.save_flax_msgpack(tree, tmp_path / "vae" / "diffusion_flax_model.msgpack")
    e2, d2 = CK.load_vae_flax(tmp_path / "vae", blocks, 3, 4, 1)
    assert list(e2) == list(enc) and all(np.array_equal(e2[k], enc[k]) for k in enc)
    assert list(d2) == list(dec) and all(np.array_equal(d2[k], dec[k]) for k in dec)
    with pytest.raises(ValueError):
        CK.load_vae_flax(tmp_path / "vae", (32, 32), 3, 4, 1)            # wrong topology: shape mismatch is reported by name
    broken = P.nest({k: v for k, v in enc.items() if k != "quant_conv/bias"})
    with pytest.raises(KeyError):
        CK.select_params(broken, P.vae_encoder_spec(blocks, 3, 4, 1))
