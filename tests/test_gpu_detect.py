"""GPU parity of rcr::detection_model (load / detect / save) against the oracle and the committed
reference-HOG goldens."""
import numpy as np
import pytest

import synth
from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_model_load_save_round_trip(sd, golden, tmp_path):
    m = sd.load_detection_model(golden.model_path)
    assert m.num_levels == 4 and m.num_landmarks == 22 and m.landmark_ids[:3] == ["9", "31", "32"]
    out = tmp_path / "rt.bin"
    sd.save_detection_model(m, str(out))
    assert out.read_bytes() == open(golden.model_path, "rb").read()
    with pytest.raises(RuntimeError):
        sd.load_detection_model(str(tmp_path / "missing.bin"))


def test_detect_example_frames_match_reference_goldens(sd, oracle, golden):
    """detect(image, facebox) on the reference's 5 annotated frames == the oracle driven by the reference's hog.c."""
    m = sd.load_detection_model(golden.model_path)
    om = oracle.Model(golden.model_path)
    for i in range(5):
        gray, box = golden.examples[f"gray{i}"], golden.examples["boxes"][i]
        lm = m.detect(gray, box)
        ref = golden.detect[f"landmarks{i}"]
        d = np.max(np.abs(lm - ref))
        print(f"frame {i}: max |landmark diff| = {d:.3e} px")
        assert d <= 1e-4 * np.max(np.abs(ref))
        # detect(image, initialisation) entry (model.hpp:147-157)
        init = oracle.align_mean(om.mean, box)
        assert np.max(np.abs(m.detect(gray, init) - ref)) <= 1e-4 * np.max(np.abs(ref))


def _rounding_margin(oracle, om, image, x):
    """Smallest distance to a rounding tie (crop centres: cvRound; half patch size: std::round of rel * IED / 2) met by the
    oracle's cascade on this face."""
    margin = 1.0
    x = np.asarray(x, dtype=np.float32).copy()
    for level in range(om.num_levels):
        hp = om.hog_params[level]
        margin = min(margin, float(np.min(np.abs(np.abs(x - np.floor(x)) - 0.5))))
        ied = oracle.get_ied(x, om.right_idx, om.left_idx)
        h = float(np.float32(hp.relative_patch_size)) * ied / 2.0
        margin = min(margin, abs(abs(h - np.floor(h)) - 0.5))
        feat = oracle.hog_transform(image, x, hp, om.right_idx, om.left_idx)
        upd = oracle.predict(feat.reshape(1, -1), om.weights[level]).ravel()
        x = (x - upd * np.float32(1.0 / np.float32(1.0 / ied))).astype(np.float32)
    return margin


def test_detect_batch_host_and_device_paths(sd, oracle, golden):
    """End-to-end cascade on seeded synthetic frames (25 % of the boxes hang over the border).

    Every cascade level rounds the landmark coordinates to integer crop centres (cvRound) and the IED to an
    integer half patch size, so a 1e-6 px difference in a landmark that sits on a rounding boundary moves a
    whole patch by one pixel: the end-to-end comparison is therefore made per face -- the bulk must agree to
    1e-4, a boundary flip may move a face by a fraction of a pixel -- and the strict 1e-4 parity is asserted
    level by level from identical inputs (teacher forcing) below."""
    import torch
    m = sd.load_detection_model(golden.model_path)
    om = oracle.Model(golden.model_path)
    images = synth.smooth_images(24, 240, 320, seed=1234)
    boxes = synth.face_boxes(24, 240, 320, seed=1234, border_fraction=0.25)
    ref = om.detect_batch(images, boxes, threads=8)
    got = m.detect_batch(images, boxes)
    per_face = np.max(np.abs(got - ref), axis=1) / np.max(np.abs(ref))
    print("batch host path: per-face rel err", np.sort(per_face)[::-1][:5], "faces within 1e-4:", int((per_face <= 1e-4).sum()), "of", len(per_face))
    # every face must agree to 1e-4 -- unless it is PROVEN to sit on a rounding boundary: replaying the oracle's cascade level by
    # level, some crop centre (cvRound) or half patch size (std::round) of that face lies within 5e-5 px of a tie, where a
    # difference in the seventh digit legitimately moves a whole patch by one pixel
    flipped = [i for i in range(len(per_face)) if per_face[i] > 1e-4]
    for i in flipped:
        near = _rounding_margin(oracle, om, images[i], np.asarray(oracle.align_mean(om.mean, boxes[i])))
        print(f"face {i}: rel err {per_face[i]:.2e}, closest rounding margin over the cascade {near:.2e}")
        assert near <= 5e-5, f"face {i} differs by {per_face[i]:.2e} without sitting on a rounding boundary"
    assert len(flipped) <= 2
    assert np.max(np.abs(got - ref)) <= 1.0          # a flipped rounding moves a landmark by well under a pixel
    x0 = np.stack([sd.align_mean(m.get_mean(), b) for b in boxes])
    assert np.array_equal(x0, np.stack([oracle.align_mean(om.mean, b) for b in boxes]))
    dev = m.detect_batch_device(torch.from_numpy(images).cuda(), torch.from_numpy(x0).cuda()).cpu().numpy()
    assert np.array_equal(dev, got)
    # determinism: same inputs -> bit-identical landmarks
    assert np.array_equal(m.detect_batch(images, boxes), got)
    # pinned host frames take the region-of-interest upload route (zero-copy gather of the face neighbourhood);
    # it must give bit-identical landmarks, including for the boxes that hang over the frame border
    pinned = torch.from_numpy(images).pin_memory()
    fb0 = m.ctx.roi_fallbacks()
    roi = m.detect_batch(pinned.numpy(), boxes)
    assert np.array_equal(roi, got)
    print("ROI route: fallbacks", m.ctx.roi_fallbacks() - fb0, "of", len(boxes))


def test_cascade_levels_teacher_forced(sd, oracle, golden):
    """One cascade level at a time from the ORACLE's landmarks: features and the updated landmarks must
    agree to 1e-4 (no rounding-boundary amplification possible inside a single level)."""
    import torch
    om = oracle.Model(golden.model_path)
    images = synth.smooth_images(12, 240, 320, seed=77)
    boxes = synth.face_boxes(12, 240, 320, seed=77, border_fraction=0.25)
    cur = np.stack([oracle.align_mean(om.mean, b) for b in boxes])
    hps = [sd.HoGParam(p.variant, p.num_cells, p.cell_size, p.num_bins, p.relative_patch_size) for p in om.hog_params]
    ht = sd.HogTransform(images, hps, om.landmark_ids, om.right_ids, om.left_ids)
    norm = sd.InterEyeDistanceNormalisation(om.landmark_ids, om.right_ids, om.left_ids)
    for level in range(om.num_levels):
        A_ref = oracle.hog_transform_batch(images, cur, om.hog_params[level], om.right_idx, om.left_idx, threads=8)
        A = ht(cur, level).cpu().numpy()
        assert rel_err(A, A_ref) <= 1e-4
        reg = sd.LinearRegressor()
        reg._ctx()
        reg.x = torch.from_numpy(om.weights[level]).cuda()
        # one level of test(): x - (A X) * IED(x)
        one = sd.SupervisedDescentOptimiser([reg], norm)
        ht1 = sd.HogTransform(images, [hps[level]], om.landmark_ids, om.right_ids, om.left_ids)
        got = one.test(cur, None, ht1).cpu().numpy()
        upd = oracle.predict(A_ref, om.weights[level])
        ref = np.empty_like(cur)
        for i in range(cur.shape[0]):
            ied = oracle.get_ied(cur[i], om.right_idx, om.left_idx)
            n = np.float32(1.0 / ied)
            ref[i] = cur[i] - upd[i] * (np.float32(1.0) / n)
        e = rel_err(got, ref)
        print(f"level {level}: features rel err {rel_err(A, A_ref):.2e}, updated landmarks rel err {e:.2e}")
        assert e <= 1e-4
        cur = ref


def test_trained_model_round_trip_through_file(sd, oracle, golden, tmp_path):
    """A model assembled from parts (detection_model ctor) saves to a file the oracle reader accepts."""
    om = oracle.Model(golden.model_path)
    sdo = sd.SupervisedDescentOptimiser([sd.LinearRegressor(sd.Regulariser(sd.RegularisationType.MatrixNorm, 1.5, False)) for _ in range(4)])
    import torch
    for k, r in enumerate(sdo.regressors):
        r.x = torch.from_numpy(om.weights[k]).cuda()
    hps = [sd.HoGParam(p.variant, p.num_cells, p.cell_size, p.num_bins, p.relative_patch_size) for p in om.hog_params]
    m = sd.detection_model.from_parts(sdo, om.mean, om.landmark_ids, hps, om.right_ids, om.left_ids)
    out = tmp_path / "again.bin"
    m.save(str(out))
    assert out.read_bytes() == open(golden.model_path, "rb").read()
