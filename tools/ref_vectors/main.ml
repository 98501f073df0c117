(* Golden vectors from the reference itself — the recipe that pins the oracle (VERDICT r1, item 9).

   Runs the UNMODIFIED functors  Groth16.Make / Pinocchio.Make  (src/groth16/groth16.ml,
   src/pinocchio/pinocchio.ml) over Curve.Bls12_381 arithmetic (opam bls12-381 6.1.0) on small
   hand-built circuits and prints ONE JSON document with everything a byte-for-byte comparison needs:
   the gates in the order QAP.build assigned them to the domain points 0..n-1 (QAP.ml:18-94), every
   scalar the protocol drew, the witness, and pkey / vkey / proof through the reference's own
   [@@deriving yojson] converters.

   Randomness.  The reference draws with [Fr.gen rng] (curve.ml:136) and the protocols are functors
   over the curve (groth16.mli:3, pinocchio.mli:3), so the curve handed to them here is
   Curve.Bls12_381 with ONE member replaced: [Fr.gen] pops the next scalar of a list filled in
   below.  Trapdoor, (r, s) and (dv, dw, dy) are therefore known whatever Random.State does, and
   they are consumed in the reference's own order (groth16.ml:51-55,124-125; pinocchio.ml:83-91,
   428-430) because it is the reference's code that asks for them.

   tests/test_cpu_reference_vectors.py (B200 repository) replays every case through oracle/zk.py and
   compares: gate order, pkey / vkey points, proof bytes. *)

module L = Bls12_381                  (* the opam library itself: raw byte encodings at the boundary *)
open Zukelang
module B = Curve.Bls12_381            (* the reference's curve module (curve.ml:77-220) over it *)

let feed : B.Fr.t list ref = ref []
let drawn : B.Fr.t list ref = ref []          (* most recent first *)

module C = struct
  module Fr = struct
    include B.Fr
    let gen (_ : Gen.rng) =
      match !feed with
      | x :: rest -> feed := rest; drawn := x :: !drawn; x
      | [] -> failwith "ref_vectors: the scalar feed is empty"
  end
  module G1 = B.G1
  module G2 = B.G2
  module GT = B.GT
  module Pairing = B.Pairing
end

module Fr = C.Fr
module Circuit = Circuit.Make (Fr)
module QAP = QAP.Make (Fr)
module G16 = Groth16.Make (C)
module Pin = Pinocchio.Make (C)

(* deterministic scalars: x <- x * k + c mod r, starting from the case's seed *)
let scalars seed n =
  let k = Z.of_string "0x9e3779b97f4a7c15f39cc0605cedc8341082276bf3a8b2c12545f4910f6c7d3b" in
  let rec go x i acc =
    if i = n then List.rev acc
    else
      let x = Z.(erem ((x * k) + of_int 0x5A554B45) Fr.order) in
      go x (i + 1) (Fr.of_z x :: acc)
  in
  go (Z.of_int seed) 0 []

(* all scalars drawn since the last call, in draw order *)
let take_drawn () = let l = List.rev !drawn in drawn := []; l

(* ---- JSON helpers -------------------------------------------------------------------------- *)
let fr (x : Fr.t) : Yojson.Safe.t = Fr.yojson_of_t x            (* decimal string, curve.ml:139 *)
let var ((s, n) : Var.t) : Yojson.Safe.t = `List [ `String s; `Int n ]
let vars s : Yojson.Safe.t = `List (List.map var (Var.Set.elements s))
let affine (a : Circuit.Affine.t) : Yojson.Safe.t =
  `List (List.map (fun (v, c) -> `List [ var v; fr c ]) (Var.Map.bindings a))
let gate (g : Circuit.Gate.t) : Yojson.Safe.t =
  `Assoc [ ("lhs", affine g.Circuit.Gate.lhs); ("l", affine g.Circuit.Gate.l); ("r", affine g.Circuit.Gate.r) ]
let poly (p : Fr.t list) : Yojson.Safe.t = `List (List.map fr p)
let poly_map m : Yojson.Safe.t =
  `List (List.map (fun (v, p) -> `List [ var v; poly p ]) (Var.Map.bindings m))
let hex_of_json (j : Yojson.Safe.t) : Yojson.Safe.t =
  (* the document as the reference writes it (Yojson.Safe.to_string: raw point bytes inside JSON
     strings), hex-encoded so that no transport re-encodes it *)
  `String (Hex.show (Hex.of_string (Yojson.Safe.to_string j)))

(* ---- circuits ------------------------------------------------------------------------------ *)
let one = Circuit.one
let v name n : Var.t = (name, n)
let aff l : Circuit.Affine.t = Var.Map.of_list (List.map (fun (x, c) -> (x, Fr.of_int c)) l)
let mk lhs l r : Circuit.Gate.t = { Circuit.Gate.lhs = aff lhs; l = aff l; r = aff r }

(* README x*x*x + x + 3 as Comp.compile shapes it (config 1): ca = x x; cb = ca x; out = (cb + x + 3)·1 *)
let cubic () =
  let x = v "input" 2 and ca = v "_tmp" 3 and cb = v "_tmp" 4 and out = v "v" 5 in
  let gates = Circuit.Gate.Set.of_list
      [ mk [ (ca, 1) ] [ (x, 1) ] [ (x, 1) ];
        mk [ (cb, 1) ] [ (ca, 1) ] [ (x, 1) ];
        mk [ (out, 1) ] [ (cb, 1); (x, 1); (one, 3) ] [ (one, 1) ] ] in
  let circuit : Circuit.t =
    { Circuit.gates; inputs_public = Var.Set.singleton one; outputs = Var.Set.singleton out;
      mids = Var.Set.of_list [ x; ca; cb ] } in
  let sol xv =
    let x2 = Fr.(xv * xv) in
    let x3 = Fr.(x2 * xv) in
    Var.Map.of_list [ (one, Fr.one); (x, xv); (ca, x2); (cb, x3); (out, Fr.(x3 + xv + of_int 3)) ] in
  (circuit, sol)

(* multiply chain c_{i+1} = c_i * x with n gates (configs 3 / 5 in small): exercises Gate.compare on
   many gates of the same shape, a two-term right factor and a public input *)
let chain n () =
  let x = v "input" 2 in
  let c i = if i = 0 then x else v "_tmp" (2 + i) in
  let out = v "v" (3 + n) in
  let body = List.init (n - 1) (fun i -> mk [ (c (i + 1), 1) ] [ (c i, 1) ] [ (x, 1) ]) in
  let last = mk [ (out, 1) ] [ (c (n - 1), 1); (one, 5) ] [ (one, 1); (x, 2) ] in
  let gates = Circuit.Gate.Set.of_list (last :: body) in
  let circuit : Circuit.t =
    { Circuit.gates; inputs_public = Var.Set.of_list [ one; x ]; outputs = Var.Set.singleton out;
      mids = Var.Set.of_list (List.init (n - 1) (fun i -> c (i + 1))) } in
  let sol xv =
    let rec go i cur acc =
      if i = n then (cur, acc) else let nx = Fr.(cur * xv) in go (i + 1) nx ((c i, nx) :: acc) in
    let last_c, mids = go 1 xv [] in
    Var.Map.of_list
      ([ (one, Fr.one); (x, xv);
         (out, Fr.((last_c + of_int 5) * (Fr.one + (xv * of_int 2)))) ] @ mids) in
  (circuit, sol)

(* ---- one case per protocol ----------------------------------------------------------------- *)
let describe name (circuit : Circuit.t) (qap : QAP.t) assignment sol : (string * Yojson.Safe.t) list =
  [ ("name", `String name);
    (* (domain point, gate): the order Gate.Set / QAP.build put the gates in *)
    ("gates", `List (List.map (fun (i, g) -> `List [ `Int i; gate g ]) assignment));
    ("inputs_public", vars circuit.Circuit.inputs_public);
    ("outputs", vars circuit.Circuit.outputs);
    ("mids", vars circuit.Circuit.mids);
    ("qap", `Assoc [ ("v", poly_map qap.QAP.v); ("w", poly_map qap.QAP.w); ("y", poly_map qap.QAP.y);
                     ("target", poly qap.QAP.target) ]);
    ("sol", `List (List.map (fun (k, x) -> `List [ var k; fr x ]) (Var.Map.bindings sol))) ]

let dummy_rng = Random.State.make [| 0 |]

let groth16_case name seed build =
  let circuit, witness = build () in
  let qap, assignment = QAP.build circuit.Circuit.gates in
  feed := scalars seed 8;
  drawn := [];
  let xv = Fr.gen dummy_rng in
  ignore (take_drawn ());
  let sol = witness xv in
  let pkey, vkey = G16.keygen dummy_rng circuit qap in
  let trapdoor = take_drawn () in                                  (* a, b, gm, d, t *)
  let proof = G16.prove dummy_rng qap pkey sol in
  let rs = take_drawn () in                                        (* r, s *)
  let public = Var.Map.filter (fun k _ -> not (Var.Set.mem k circuit.Circuit.mids)) sol in
  let ok = G16.verify public vkey proof in
  `Assoc (describe name circuit qap assignment sol
          @ [ ("trapdoor", `List (List.map fr trapdoor)); ("rs", `List (List.map fr rs));
              ("pkey_hex", hex_of_json (G16.yojson_of_pkey pkey));
              ("vkey_hex", hex_of_json (G16.yojson_of_vkey vkey));
              ("proof_hex", hex_of_json (G16.yojson_of_proof proof));
              ("verified", `Bool ok) ])

(* NonZK and ZK are two Protocol.S instances with their own abstract key types (pinocchio.mli:3-15),
   so each generates its key — from the same scalar feed, hence the same key. *)
let pinocchio_case name seed build =
  let circuit, witness = build () in
  let qap, assignment = QAP.build circuit.Circuit.gates in
  let start () =
    feed := scalars seed 16;
    drawn := [];
    let xv = Fr.gen dummy_rng in
    ignore (take_drawn ());
    xv in
  let sol = witness (start ()) in
  let public = Var.Map.filter (fun k _ -> not (Var.Set.mem k circuit.Circuit.mids)) sol in
  let pkey, vkey = Pin.ZK.keygen dummy_rng circuit qap in
  let trapdoor = take_drawn () in                                  (* rv, rw, s, av, aw, ay, b, gm *)
  let proof_zk = Pin.ZK.prove dummy_rng qap pkey sol in
  let d = take_drawn () in                                         (* dv, dw, dy *)
  let ok_zk = Pin.ZK.verify public vkey proof_zk in
  ignore (start ());
  let pkey_n, vkey_n = Pin.NonZK.keygen dummy_rng circuit qap in
  ignore (take_drawn ());
  let proof_nonzk = Pin.NonZK.prove dummy_rng qap pkey_n sol in
  let nonzk_draws = take_drawn () in                               (* expected: none *)
  let ok_n = Pin.NonZK.verify public vkey_n proof_nonzk in
  `Assoc (describe name circuit qap assignment sol
          @ [ ("trapdoor", `List (List.map fr trapdoor));
              ("nonzk_draws", `List (List.map fr nonzk_draws));
              ("d", `List (List.map fr d));
              ("pkey_hex", hex_of_json (Pin.ZK.yojson_of_pkey pkey));
              ("vkey_hex", hex_of_json (Pin.ZK.yojson_of_vkey vkey));
              ("proof_nonzk_hex", hex_of_json (Pin.NonZK.yojson_of_proof proof_nonzk));
              ("proof_zk_hex", hex_of_json (Pin.ZK.yojson_of_proof proof_zk));
              ("verified", `Bool (ok_zk && ok_n)) ])

(* raw encodings at the bls12-381 boundary (curve.ml:139-140,199,208): Fr.to_bytes endianness, the
   compressed sign rule on both groups, the identity *)
let boundary () =
  let h b = `String (Hex.show (Hex.of_bytes b)) in
  let ks = scalars 0xB0 4 in
  `Assoc
    [ ("fr", `List (List.map (fun k -> `List [ fr k; h (L.Fr.to_bytes k) ]) ks));
      ("g1", `List (List.map (fun k ->
           let p = B.G1.of_Fr k in
           `List [ fr k; h (L.G1.to_bytes p); h (L.G1.to_compressed_bytes p) ]) ks));
      ("g2", `List (List.map (fun k ->
           let p = B.G2.of_Fr k in
           `List [ fr k; h (L.G2.to_bytes p); h (L.G2.to_compressed_bytes p) ]) ks));
      ("g1_zero", `List [ h (L.G1.to_bytes B.G1.zero); h (L.G1.to_compressed_bytes B.G1.zero) ]);
      ("g2_zero", `List [ h (L.G2.to_bytes B.G2.zero); h (L.G2.to_compressed_bytes B.G2.zero) ]) ]

let () =
  let doc =
    `Assoc
      [ ("source", `String "camlspotter/zukelang, bls12-381 6.1.0; tools/ref_vectors/main.ml");
        ("boundary", boundary ());
        ("groth16", `List [ groth16_case "cubic" 1 cubic; groth16_case "chain8" 2 (chain 8) ]);
        ("pinocchio", `List [ pinocchio_case "cubic" 3 cubic; pinocchio_case "chain8" 4 (chain 8) ]) ]
  in
  print_string (Yojson.Safe.pretty_to_string doc);
  print_newline ()
