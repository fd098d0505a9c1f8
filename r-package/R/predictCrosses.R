# Drop-in replacement of genomicMateSelectR::predictCrosses (R/gmsFunctions.R:650-880 of the reference) on top of
# libgms_b200.so: ONE device pass per call (DirDom's two predCrossVars calls of :704-723 are one pass with three
# components), cross means (predCrossMeans, R/predCrossVar.R:336-399) and the whole tidy table (TGV = BV + DD or A + D,
# selection-index mean and variance, predSD, usefulness; :764-871) computed on the device by gmsr_tidy.
# This R code has never been run (no R in the build image); the Python twin is predictCrosses() in
# genomicmateselectr_b200/predcrossvar.py, the .Call entry points are executed by tests/test_gpu_rshim.py.

.gms_efflist <- function(snpeffs, col) {
  if (!col %in% colnames(snpeffs)) stop(paste0("snpeffs lacks column ", col))
  # snpeffs$<col> is a list of m x 1 matrices with rownames = SNP ids; predCrossVars wants 1 x m (:661-663)
  stats::setNames(lapply(snpeffs[[col]], t), snpeffs$Trait)
}

predictCrosses <- function(modelType, stdSelInt = 2.67, selInd, SIwts = NULL, CrossesToPredict, snpeffs, dosages,
                           haploMat, recombFreqMat, ncores = 1, nBLASthreads = NULL,
                           predTheMeans = TRUE, predTheVars = TRUE, session = NULL) {
  if (!modelType %in% c("A", "AD", "DirDom")) stop("modelType must be 'A', 'AD' or 'DirDom'")
  if (is.null(session)) {
    parents <- union(CrossesToPredict$sireID, CrossesToPredict$damID)
    session <- gms_session(haploMat, recombFreqMat, parents)
    on.exit(lapply(session$handles, function(h) .Call("gmsr_close", h)))
  }
  asub <- .gms_efflist(snpeffs, "allelesubsnpeff")
  traits <- names(asub)
  model <- match(modelType, c("A", "AD", "DirDom")) - 1L
  print("Predicting cross variance parameters")
  starttime <- proc.time()[3]
  res <- switch(modelType,
    A = .gms_run(session, CrossesToPredict, 0L, asub),
    AD = .gms_run(session, CrossesToPredict, 1L, asub, .gms_efflist(snpeffs, "domdevsnpeff")),
    DirDom = .gms_run(session, CrossesToPredict, 2L, .gms_efflist(snpeffs, "addsnpeff"),
                      .gms_efflist(snpeffs, "domsnpeff"), asub))
  # the slab is still in HBM: means, TGV, SELIND, predSD and usefulness in one more device pass
  w <- if (selInd) { x <- SIwts[traits]; x[is.na(x)] <- 0; as.double(x) } else NULL
  tab <- .Call("gmsr_tidy", session$handle, model, length(traits), nrow(CrossesToPredict), w, as.double(stdSelInt))
  print(paste0("Done predicting fam vars. Took ", round((proc.time()[3] - starttime) / 60, 2), " mins for ",
               sum(res$nseg > 0), " crosses"))
  keep <- which(res$nseg > 0)                                     # zero-segregation crosses vanish (R/predCrossVar.R:156-160)
  ofs <- if (modelType == "A") "BV" else c("BV", "TGV")
  slots <- c(if (selInd) "SELIND", traits)
  # reference row order (:777-799, :825-853): all BV rows, then all TGV rows; per cross SELIND first, then the traits
  tidy <- do.call(rbind, lapply(seq_along(ofs), function(oi) {
    v <- tab[, , oi, keep, drop = FALSE]                          # 4 x nslots x 1 x nkeep
    tibble::tibble(sireID = rep(CrossesToPredict$sireID[keep], each = length(slots)),
                   damID = rep(CrossesToPredict$damID[keep], each = length(slots)),
                   Nsegsnps = rep(res$nseg[keep], each = length(slots)),
                   predOf = ofs[oi], Trait = rep(slots, times = length(keep)),
                   predMean = as.vector(v[1, , 1, ]), predVar = as.vector(v[2, , 1, ]),
                   predSD = as.vector(v[3, , 1, ]), predUsefulness = as.vector(v[4, , 1, ]))
  }))
  if (!predTheMeans) tidy <- tidy[, c("sireID", "damID", "Nsegsnps", "predOf", "Trait", "predVar", "predSD")]
  if (!predTheVars) tidy <- tidy[, c("sireID", "damID", "predOf", "Trait", "predMean")]
  comps <- list(A = "VarBV", AD = c("VarBV", "VarDD"), DirDom = c("VarA", "VarD", "VarBV"))[[modelType]]
  pairs <- rbind(cbind(traits, traits), if (length(traits) > 1) t(combn(traits, 2)))
  rawvars <- tibble::tibble(sireID = rep(CrossesToPredict$sireID[keep], each = nrow(pairs) * length(comps)),
                            damID = rep(CrossesToPredict$damID[keep], each = nrow(pairs) * length(comps)),
                            Nsegsnps = rep(res$nseg[keep], each = nrow(pairs) * length(comps)),
                            Trait1 = rep(rep(pairs[, 1], each = length(comps)), times = length(keep)),
                            Trait2 = rep(rep(pairs[, 2], each = length(comps)), times = length(keep)),
                            predOf = rep(comps, times = nrow(pairs) * length(keep)),
                            predVar = as.vector(res$out[, , keep]))
  tibble::tibble(tidyPreds = list(tidy), rawPreds = list(list(predVars = list(rawvars))))
}
