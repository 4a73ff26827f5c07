from nonstationary_precip_b200.utils.dataprep import kmeans_inducing_points  # noqa: F401
