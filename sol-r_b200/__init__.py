"""solr_b200 — B200-native engine for Sol-R's ray-propagation hot path."""
