"""CoFusion::processFrame sequencing (Core/CoFusion.cpp:171-545) driven through the CPU oracle --
the checker for cofusion_b200.CoFusion.  TEST INFRASTRUCTURE."""
import numpy as np

import orc


class OraclePipeline:
    def __init__(self, W, H, K, max_surfels=1 << 20, conf_global=10.0, depth_cutoff=5.0, max_depth=20.0,
                 icp_weight=10.0, time_delta=200, outlier_coeff=3.0, predict_before_fuse=True):
        self.W, self.H, self.K = W, H, K
        # CoFusion.cpp:347 renders a prediction between tracking and fusing that only the loop closure reads
        # (out of scope); the product skips it (predictBeforeFuse = 0).  Results do not depend on it.
        self.predict_before_fuse = predict_before_fuse
        self.map = orc.OrcMap(W, H, K, max_surfels)
        self.odom = orc.OrcOdometry(W, H, K)
        self.pose = np.eye(4, dtype=np.float32)
        self.last_pose = np.eye(4, dtype=np.float32)
        self.tick = 1
        self.conf = conf_global
        self.depth_cutoff, self.max_depth = depth_cutoff, max_depth
        self.icp_weight, self.time_delta, self.outlier = icp_weight, time_delta, outlier_coeff
        self.stats = None

    def predict(self, rgb, df):
        self.map.combined_predict(self.pose, self.max_depth, self.conf, self.tick, self.tick, self.time_delta)
        self.map.fill_in(rgb, df, False, False)

    def process_frame(self, rgb, depth, mask=None):
        mask = np.zeros((self.H, self.W), np.uint8) if mask is None else mask
        df = orc.bilateral(depth, self.depth_cutoff)
        if self.tick == 1:
            self.map.initialise(rgb, depth, df, self.tick, self.max_depth)
            self.odom.init_first_rgb(rgb)
        else:
            m = self.map
            if m.requires_fill_in():  # Model::initICP with doFillIn (Model.cpp:354-356)
                v4, n4, img = m.view(9), m.view(10), m.view(8)
            else:
                v4, n4, img = m.view(5), m.view(6), m.view(4)
            self.last_pose = self.pose.copy()
            self.odom.init_model(v4, n4, img, self.pose)
            self.odom.init_frame(df, rgb, self.max_depth)
            self.pose, self.stats, _, _ = self.odom.track(self.pose, icp_weight=self.icp_weight)
            if self.predict_before_fuse:
                self.predict(rgb, df)
            w = orc.OrcMap.fusion_weight(self.pose, self.last_pose, 1.0)
            m.predict_indices(self.pose, self.tick, self.max_depth, self.time_delta)
            m.fuse(self.pose, self.tick, rgb, mask, depth, df, self.max_depth, w, 0)
            m.predict_indices(self.pose, self.tick, self.max_depth, self.time_delta)
            m.clean(self.pose, self.tick, self.conf, self.time_delta, df, mask, 0, self.outlier)
        self.predict(rgb, df)
        self.tick += 1


class _OrcModel:
    def __init__(self, W, H, K, model_id, conf, max_surfels, fill_in):
        self.map = orc.OrcMap(W, H, K, max_surfels)
        self.odom = orc.OrcOdometry(W, H, K)
        self.pose = np.eye(4, dtype=np.float32)
        self.last_pose = np.eye(4, dtype=np.float32)
        self.id, self.conf, self.fill_in = model_id, np.float32(conf), fill_in
        self.max_depth = np.float32(np.finfo(np.float32).max)
        self.icp_error = np.zeros((H, W), np.float32)
        self.stats = None


class OracleCoFusion:
    """CoFusion::processFrame with enableMultipleModels (CoFusion.cpp:171-524, :227-299 for the
    segmentation-driven model management), on the CPU oracle."""

    def __init__(self, W, H, K, max_surfels=1 << 20, conf_global=10.0, conf_object=0.01, depth_cutoff=5.0,
                 max_depth=20.0, icp_weight=10.0, time_delta=200, outlier_coeff=3.0, spawn_offset=20,
                 seg_params=None, multiple_models=True):
        self.W, self.H, self.K, self.max_surfels = W, H, K, max_surfels
        self.models = [_OrcModel(W, H, K, 0, conf_global, max_surfels, True)]
        self.inactive = []
        self.tick = 1
        self.conf_object = conf_object
        self.depth_cutoff, self.max_depth = depth_cutoff, max_depth
        self.icp_weight, self.time_delta, self.outlier = icp_weight, time_delta, outlier_coeff
        self.multiple = multiple_models
        self.model_spawn_offset, self.spawn_offset, self.next_id = spawn_offset, 0, 1
        self.seg_params = seg_params or orc.OrcSegParams.default()
        self.mask = np.zeros((H, W), np.uint8)
        self.last_seg = None  # (mds, has_new, spawned_id, deactivated)

    def _take_next_id(self):  # CoFusion::getNextModelID(true)
        nxt = self.next_id
        while True:
            self.next_id = (self.next_id + 1) & 255
            if all(m.id != self.next_id for m in self.models):
                break
        return nxt

    @staticmethod
    def _seg_max_depth(md):  # getMaxDepth lambda (CoFusion.cpp:228): float + float * double
        return np.float32(np.float64(np.float32(md["depthMean"])) + np.float64(np.float32(md["depthStd"])) * 1.2)

    def predict(self, rgb, df):
        for m in self.models:
            m.map.combined_predict(m.pose, self.max_depth, m.conf, self.tick, self.tick, self.time_delta)
            if m.fill_in:
                m.map.fill_in(rgb, df, False, False)

    def _segment(self, rgb, depth, df):
        if self.spawn_offset < self.model_spawn_offset:
            self.spawn_offset += 1
        n = len(self.models)
        allow_new = self.spawn_offset >= self.model_spawn_offset and n < 15
        owners = list(self.models)
        seg, mds, has_new, _, _, _ = orc.segment_crf(rgb, depth, [m.id for m in self.models],
                                                     [m.icp_error for m in self.models],
                                                     [m.map.view(5) for m in self.models], self.next_id, allow_new,
                                                     self.seg_params)
        self.mask = seg
        spawned, deactivated = -1, 0
        new_model = None
        if has_new:
            new_model = _OrcModel(self.W, self.H, self.K, self._take_next_id(), self.conf_object, self.max_surfels, False)
            new_model.odom.init_first_rgb(rgb)
            self.spawn_offset = 0
            new_model.max_depth = self._seg_max_depth(mds[-1])
            spawned = new_model.id
        for i in range(1, len(self.models)):
            self.models[i].max_depth = self._seg_max_depth(mds[i])
        if has_new:
            m = new_model
            m.map.predict_indices(m.pose, self.tick, self.max_depth, self.time_delta)
            w = orc.OrcMap.fusion_weight(m.pose, m.last_pose, 100.0)
            m.map.fuse(m.pose, self.tick, rgb, seg, depth, df, min(np.float32(self.max_depth), m.max_depth), w, m.id)
            m.map.clean(m.pose, self.tick, m.conf, self.time_delta, df, seg, m.id, self.outlier)
            self.models.append(m)
        for k in range(min(len(mds), n)):
            if mds[k]["superPixelCount"] <= 0 and mds[k]["id"] != 0:
                self.models.remove(owners[k])
                self.inactive.append(owners[k])
                deactivated += 1
        for i in range(1, min(len(self.models), len(mds))):  # positional indexing after the list changed
            m = self.models[i]
            m.conf = np.float32(min(max(m.conf, np.float32(mds[i]["avgConfidence"])), np.float32(9.0)))
        self.last_seg = (mds, has_new, spawned, deactivated)

    def process_frame(self, rgb, depth):
        df = orc.bilateral(depth, self.depth_cutoff)
        if self.tick == 1:
            g = self.models[0]
            g.map.initialise(rgb, depth, df, self.tick, self.max_depth)
            g.odom.init_first_rgb(rgb)
        else:
            for m in self.models:
                if m.fill_in and m.map.requires_fill_in():
                    v4, n4, img = m.map.view(9), m.map.view(10), m.map.view(8)
                else:
                    v4, n4, img = m.map.view(5), m.map.view(6), m.map.view(4)
                m.last_pose = m.pose.copy()
                m.odom.init_model(v4, n4, img, m.pose)
                m.odom.init_frame(df, rgb, self.max_depth)
                m.pose, m.stats, m.icp_error, _ = m.odom.track(m.pose, icp_weight=self.icp_weight, want_error=True)
            if self.multiple:
                self._segment(rgb, depth, df)
            ms = self.models
            for m in ms:
                m.map.predict_indices(m.pose, self.tick, self.max_depth, self.time_delta)
            for m in ms:
                w = orc.OrcMap.fusion_weight(m.pose, m.last_pose, 1.0)
                m.map.fuse(m.pose, self.tick, rgb, self.mask, depth, df, min(np.float32(self.max_depth), m.max_depth), w,
                           m.id)
            for m in ms:
                m.map.predict_indices(m.pose, self.tick, self.max_depth, self.time_delta)
            for m in ms:
                m.map.clean(m.pose, self.tick, m.conf, self.time_delta, df, self.mask, m.id, self.outlier)
        self.predict(rgb, df)
        self.tick += 1
